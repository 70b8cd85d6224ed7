// Forwarding header of the Ceres-compatibility shim: everything lives in ceres/ceres.h.
#include "ceres/ceres.h"
