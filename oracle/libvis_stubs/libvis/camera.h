// TEST INFRASTRUCTURE (oracle build aid): libvis/camera.h pulls Sophus and the camera models; the CPU meshing code
// (surfel_meshing.{h,cc}) includes it without using anything from it.
#pragma once
#include <algorithm>
#include "libvis/eigen.h"
#include "libvis/logging.h"
