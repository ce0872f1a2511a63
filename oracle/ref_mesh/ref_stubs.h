#pragma once
#include "mpi.h"
#ifndef MPI_UNSIGNED_SHORT
#define MPI_UNSIGNED_SHORT 15
#define MPI_LONG_DOUBLE 16
#endif
// user data hooks that ampsConfig.pl injects from the physical model headers: empty in this stand-alone build
class cInternalSphericalData_UserDefined {};
class cInternalCircleData_UserDefined {};
class cInternalSphere1DData_UserDefined {};
class cInternalRotationBodyData_UserDefined {};
class cInternalNastranSurfaceData_UserDefined {};
