// Stand-in: feature_tracker.h includes it, the event front-end never uses a MEI camera.
#pragma once
#include "Camera.h"
