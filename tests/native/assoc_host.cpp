// CPU build of stemseg_b200/csrc/assoc.cuh for tests/test_assoc_cpu.py (g++ -shared; the same functions run inside
// the stitch kernel on the device).
#include "../../stemseg_b200/csrc/assoc.cuh"

extern "C" int assoc_pyset_order(const long long* values, int n, long long* out) {
    return stemseg::pyset_order(values, n, out);
}

extern "C" int assoc_lsap(const double* cost, int nr, int nc, int* rows, int* cols) {
    static stemseg::LsapScratch w;
    return stemseg::lsap_solve(cost, nr, nc, rows, cols, w);
}
