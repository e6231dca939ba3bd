// expand_math.h -- 4 bases (code bits in the low nibbles of a and b) -> 4 ASCII bytes.
// spread(n) puts bit k of the nibble at bit 0 of byte k (one multiply, no carries); the byte of code
// (b<<1 | a) is 'A' + 2a + 0x13 b - 0x0F ab  =  A 0x41, C 0x43, T 0x54, G 0x47.
#pragma once
#include <stdint.h>

#include "mdbg_common.cuh"

namespace mdbg {

MDBG_HD uint32_t expand4(uint32_t a, uint32_t b) {
    const uint32_t sa = ((a & 15u) * 0x00204081u) & 0x01010101u;
    const uint32_t sb = ((b & 15u) * 0x00204081u) & 0x01010101u;
    return 0x41414141u + 2u * sa + 0x13u * sb - 0x0Fu * (sa & sb);
}

}  // namespace mdbg
