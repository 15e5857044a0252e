// Stand-in for <opencv2/imgproc/imgproc.hpp> (oracle/_ref build only): see core/core.hpp.
#pragma once
#include "../core/core.hpp"
