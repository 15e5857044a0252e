// Stand-in for <opencv2/highgui/highgui.hpp> (oracle/_ref build only): see core/core.hpp.
#pragma once
#include "../core/core.hpp"
