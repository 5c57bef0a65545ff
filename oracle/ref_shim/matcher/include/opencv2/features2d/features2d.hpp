#include "../../../cvmini.hpp"
