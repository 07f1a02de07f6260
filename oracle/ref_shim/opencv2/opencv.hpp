/* shim: see mini_cv.h.  The real opencv2/opencv.hpp pulls in these standard headers, which
 * feature_tracker.h relies on (std::map, std::stringstream, std::setprecision). */
#include <iomanip>
#include <map>
#include <sstream>
#include <string>
#include "../mini_cv.h"
