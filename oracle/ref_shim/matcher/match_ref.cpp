// TEST INFRASTRUCTURE ONLY.
// oracle/_ref/libmatch_ref.so: the reference's own ORBmatcher.cc, compiled where it lies (REF_ORBMATCHER_CC, never
// copied) against the stand-ins of slam_shim.hpp, behind the same flat-array C signatures as the oracle's orbo_search_*
// entry points -- so a test can hand identical inputs to the oracle restatement and to the reference's code and demand
// identical outputs.  The searches whose inputs the C ABI takes as already-projected coordinates are driven with an
// identity camera (R = I, t = 0, fx = fy = 1, cx = cy = 0, depth 1), for which the reference's own projection arithmetic
// returns the given coordinates bit for bit.
#include <chrono>

#include "slam_shim.hpp"

using namespace std;   // the reference's headers lean on a using-directive leaked by the headers replaced here

#include REF_ORBMATCHER_CC

#define GLUE(name) refm_##name
#include "match_glue.inc"
