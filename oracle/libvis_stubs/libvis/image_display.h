// TEST INFRASTRUCTURE (oracle build aid): stand-in for libvis/image_display.h, which the CPU meshing code includes but only uses in
// debug paths that are not compiled here.
#pragma once
#include "libvis/eigen.h"
