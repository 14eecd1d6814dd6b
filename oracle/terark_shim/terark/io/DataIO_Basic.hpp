#pragma once
#include "../stdtypes.hpp"
#include "../util/byte_swap_impl.hpp"
