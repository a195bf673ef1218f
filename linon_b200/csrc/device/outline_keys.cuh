// outline_keys.cuh — the keys of the outline search (src/main_view.wgsl:381-468), shared by outline_search_kernel /
// outline_finalize_kernel (static_kernels.cu) and, in TEST builds only, by tests/emu/outline_driver.cpp, which compiles
// this text for the host and checks it against the oracle's eight serial scans.
//
// The reference runs eight serial scans over exponents >= 0.8.  Each becomes a maximum over 64-bit keys (larger wins,
// 0 = nothing found), so all eight are one parallel reduction:
//   "max sum, last in scan order (x outer, y inner)"  -> (sum, x, y) lexicographic maximum
//   "first found in scan order (a outer, b inner)"    -> (a, b) lexicographic minimum, stored inverted
#pragma once

#ifndef LINON_OK_HD
#define LINON_OK_HD __device__ __forceinline__
#endif

namespace linon {

constexpr unsigned long long KEY_VALID = 1ull << 63;
LINON_OK_HD unsigned long long key3(unsigned int a, unsigned int b, unsigned int c) {
    return KEY_VALID | ((unsigned long long)a << 42) | ((unsigned long long)b << 21) | (unsigned long long)c;
}
constexpr unsigned int M21 = (1u << 21) - 1u;

// the eight keys of one pixel (x, y) with exponent >= 0.8 in a uw x uh frame
LINON_OK_HD void outline_keys(unsigned int x, unsigned int y, unsigned int uw, unsigned int uh, unsigned long long (&k)[8]) {
    k[0] = key3((uw - x) + y, x, y);                   // bottom left   :385-394
    k[1] = key3(0, M21 - (uh - 1 - y), M21 - x);       // bottom middle :395-404 (rows from the bottom)
    k[2] = key3(x + y, x, y);                          // bottom right  :405-414
    k[3] = key3(0, M21 - (uw - 1 - x), M21 - y);       // middle right  :415-424
    k[4] = key3(x + (uh - y), x, y);                   // top right     :425-434
    k[5] = key3(0, M21 - y, M21 - x);                  // top middle    :435-444
    k[6] = key3((uw - x) + (uh - y), x, y);            // top left      :445-454
    k[7] = key3(0, M21 - x, M21 - y);                  // middle left   :455-464
}

// winning key of search j -> pixel position as floats (left at 0, 0 when nothing was found)
LINON_OK_HD void outline_decode(int j, unsigned long long k, int W, int H, float& px, float& py) {
    if (k & KEY_VALID) {
        const unsigned int b = (unsigned int)(k >> 21) & M21, c = (unsigned int)k & M21;
        unsigned int x, y;
        switch (j) {
            case 1: y = (unsigned int)H - 1 - (M21 - b); x = M21 - c; break;
            case 3: x = (unsigned int)W - 1 - (M21 - b); y = M21 - c; break;
            case 5: y = M21 - b; x = M21 - c; break;
            case 7: x = M21 - b; y = M21 - c; break;
            default: x = b; y = c; break;
        }
        px = (float)x; py = (float)y;
    }
}

}  // namespace linon
