// Stand-in (oracle/_ref build only): see CGAL/Kd_tree.h in this directory.
#pragma once
#include "Kd_tree.h"
