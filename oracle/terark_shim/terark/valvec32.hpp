#pragma once
#include "valvec.hpp"
