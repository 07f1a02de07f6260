// Stand-in: see bag.h.
#pragma once
#include "bag.h"
