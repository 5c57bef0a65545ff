// TEST INFRASTRUCTURE ONLY.
// oracle/_ref/libmatch_adapter.so: the adapter's reference-signature overloads (adapter/ORBmatcher.h + ORBmatcher_orbslam.inl,
// i.e. what a maintainer compiles INSTEAD of the reference's ORBmatcher.cc) against the same stand-in object graphs and
// behind the same flat-array entry points as libmatch_ref.so.  tests/test_adapter_gpu.py feeds both the same inputs.
#include <chrono>

#include "slam_shim.hpp"

using namespace std;

#define ORBB200_WITH_ORBSLAM
#include "ORBextractor.h"
#include "ORBmatcher.h"
#include "ORBmatcher_orbslam.inl"

#define GLUE_ADAPTER
#define GLUE(name) adpm_##name
#include "match_glue.inc"
