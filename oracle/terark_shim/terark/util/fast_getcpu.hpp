#pragma once
#include "../stdtypes.hpp"
#include <sched.h>
namespace terark { inline int fast_getcpu() { return sched_getcpu(); } }
