#pragma once
#include <terark/stdtypes.hpp>
