#pragma once
// stand-in for boost.predef endian detection (x86-64 / aarch64 little-endian hosts only)
#define BOOST_ENDIAN_LITTLE_BYTE 1
#define BOOST_ENDIAN_BIG_BYTE 0
