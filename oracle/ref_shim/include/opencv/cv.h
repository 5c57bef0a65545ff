// TEST INFRASTRUCTURE ONLY -- stand-in for the OpenCV header of the same name; see cvshim.hpp.
#pragma once
#include "cvshim.hpp"
