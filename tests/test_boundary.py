"""CPU: the drop-in boundary.  The C-ABI library loads without a GPU and exports every symbol
include/rtpp_gpu.h declares; the product never reaches into oracle/; without a CUDA device the
library fails loudly instead of falling back."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "rtpp_gpu.h")).read()
    names = re.findall(r"RTPP_API\s+[\w\s\*]+?\b(rtpp_\w+)\s*\(", src)
    assert len(names) >= 40
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    import rtpp_b200

    lib = rtpp_b200.lib()  # loading needs no GPU
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, f"declared in rtpp_gpu.h but not exported: {missing}"
    out = subprocess.run(["nm", "-D", "--defined-only", rtpp_b200.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (rtpp_\w+)", out))
    assert set(_declared()) <= exported
    assert lib.rtpp_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch

    import rtpp_b200

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rtpp_b200.RtppError):
        rtpp_b200.Ctx(0)
    h = ctypes.c_void_p()
    assert rtpp_b200.lib().rtpp_create(0, ctypes.byref(h)) < 0 and not h.value


def test_product_never_touches_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "rtpp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"\boracle\b", txt) and "no oracle" not in txt:
                    bad.append(os.path.join(base, f))
    # the planner layer's README-style comments may name the oracle, code must not include/load it
    for p in bad:
        txt = open(p, errors="ignore").read()
        assert not re.search(r"(#include|import|CDLL|dlopen)[^\n]*oracle", txt), p


def test_sources_build_for_sm100a_only():
    mk = open(os.path.join(ROOT, "rtpp_b200", "csrc", "Makefile")).read()
    assert "arch=compute_100a,code=sm_100a" in mk and "-lineinfo" in mk and "-fmad=false" in mk


def test_every_entry_point_rejects_a_null_handle():
    """Error behaviour of the ABI: no exceptions, no crashes — every entry point whose first parameter is a
    handle returns a negative status when it is NULL (all other arguments zero), without touching CUDA.  Run in a child process
    so that a missing guard shows up as a failed test, not as a dead test session."""
    import sys

    src = open(os.path.join(ROOT, "include", "rtpp_gpu.h")).read()
    decls = re.findall(r"RTPP_API\s+([\w\s\*]+?)\b(rtpp_\w+)\s*\(([^;]*?)\)\s*;", src, re.S)
    probes = []
    for ret, name, args in decls:
        a = [x for x in args.replace("\n", " ").split(",") if x.strip() and x.strip() != "void"]
        if ret.strip() == "int" and a and re.search(r"rtpp_\w+\s*\*", a[0]):
            probes.append((name, len(a)))
    assert len(probes) >= 80
    import rtpp_b200

    code = (
        "import ctypes as C, sys\n"
        f"lib = C.CDLL({rtpp_b200.lib_path()!r})\n"
        f"for name, n in {probes!r}:\n"
        "    fn = getattr(lib, name); fn.restype = C.c_int\n"
        "    rc = fn(*([C.c_void_p(0)] * n))\n"
        "    print(name, rc, flush=True)\n"
    )
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    seen = dict(line.split() for line in out.stdout.strip().split("\n") if line)
    assert out.returncode == 0, f"crashed after {list(seen)[-1] if seen else 'nothing'}: {out.stderr[-300:]}"
    bad = {k: v for k, v in seen.items() if int(v) >= 0}
    assert not bad and len(seen) == len(probes), bad
