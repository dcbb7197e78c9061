// Sequential pieces of the sub-clip association (OnlineChainer.associate_clusters, stemseg/inference/
// online_chainer.py:291-343) as plain host+device functions, so that the stitch never leaves the GPU and the very same
// code can be compiled with g++ and checked on the CPU against the interpreter / scipy (tests/test_assoc_cpu.py).
//
//  * lsap_solve: scipy.optimize.linear_sum_assignment (online_chainer.py:330).  scipy is a third-party dependency of
//    the reference (requirements.txt:3 pins scipy==1.10.0; this image has 1.18.1 -- same algorithm since 1.4): the
//    shortest-augmenting-path method of D. F. Crouse, "On implementing 2D rectangular assignment algorithms", IEEE
//    TAES 52(4), 2016, as scipy implements it (rows processed in order, the column scan over the `remaining` list that
//    starts REVERSED and shrinks by swap-removal, ties resolved towards an unassigned column, transposition when there
//    are more rows than columns, result rows ascending).  The reference associates clusters with IoU 0 as well
//    (cost 1.0), so ties are common and the tie-breaking has to be scipy's for bit-identical track ids.
//  * pyset_order: the order in which `list(set(sorted_ints) - {-1})` (online_chainer.py:307-308) iterates, i.e.
//    CPython's open-addressing set of small ints (hash(v) = v, hash(-1) = -2; LINEAR_PROBES 9, PERTURB_SHIFT 5, growth
//    by x4 when fill*5 >= mask*3; `a - b` copies a when len(a)//4 > len(b), else re-inserts a's entries in table
//    order).  That order decides the row / column order of the cost matrix and therefore which of several equal-cost
//    assignments the reference ends up with.  Pinned against the running interpreter (CPython 3.12) by the CPU test.
#pragma once

#include <cstdint>

#ifdef __CUDACC__
#define SS_HD __host__ __device__
#else
#define SS_HD
#endif

namespace stemseg {

constexpr int kAssocMaxSide = 72;            // labels per side of one association (max_instances 64 + outliers + slack)
constexpr int kPySetMaxTable = 512;        // sets of up to ~300 labels

// ---------------------------------------------------------------------------------------------------------------
// CPython set emulation (ints only)
// ---------------------------------------------------------------------------------------------------------------
struct PySet {
    long long key[kPySetMaxTable];      // valid where state == 1
    unsigned char state[kPySetMaxTable];   // 0 unused, 1 active, 2 dummy
    int mask, fill, used;
    int overflow;
};

SS_HD inline long long pyset_hash(long long v) { return v == -1 ? -2 : v; }

SS_HD inline void pyset_init(PySet& s) {
    s.mask = 7; s.fill = 0; s.used = 0; s.overflow = 0;
    for (int i = 0; i <= 7; ++i) s.state[i] = 0;
}

// set_insert_clean (Objects/setobject.c): table known to hold no equal key and no dummies
SS_HD inline void pyset_insert_clean(long long* key, unsigned char* state, int mask, long long v) {
    const long long h = pyset_hash(v);
    unsigned long long perturb = static_cast<unsigned long long>(h);
    unsigned long long i = static_cast<unsigned long long>(h) & static_cast<unsigned long long>(mask);
    for (;;) {
        unsigned long long e = i;
        if (state[e] == 0) { key[e] = v; state[e] = 1; return; }
        if (i + 9 <= static_cast<unsigned long long>(mask)) {
            for (int j = 0; j < 9; ++j) {
                ++e;
                if (state[e] == 0) { key[e] = v; state[e] = 1; return; }
            }
        }
        perturb >>= 5;
        i = (i * 5 + 1 + perturb) & static_cast<unsigned long long>(mask);
    }
}

SS_HD inline void pyset_resize(PySet& s, int minused) {
    int newsize = 8;
    while (newsize <= minused) newsize <<= 1;
    if (newsize > kPySetMaxTable) { s.overflow = 1; return; }
    long long old_key[kPySetMaxTable];
    unsigned char old_state[kPySetMaxTable];
    const int oldmask = s.mask;
    for (int i = 0; i <= oldmask; ++i) { old_key[i] = s.key[i]; old_state[i] = s.state[i]; }
    for (int i = 0; i < newsize; ++i) s.state[i] = 0;
    s.mask = newsize - 1;
    s.fill = s.used;
    for (int i = 0; i <= oldmask; ++i)
        if (old_state[i] == 1) pyset_insert_clean(s.key, s.state, s.mask, old_key[i]);
}

// set_add_entry for a key known to be absent or present (ints: equality == same value)
SS_HD inline void pyset_add(PySet& s, long long v) {
    const long long h = pyset_hash(v);
    const unsigned long long mask = static_cast<unsigned long long>(s.mask);
    unsigned long long perturb = static_cast<unsigned long long>(h);
    unsigned long long i = static_cast<unsigned long long>(h) & mask;
    int freeslot = -1;
    for (;;) {
        unsigned long long e = i;
        int probes = (i + 9 <= mask) ? 9 : 0;
        for (;;) {
            if (s.state[e] == 0) {
                if (freeslot >= 0) {              // found_unused_or_dummy with a dummy seen on the way
                    s.used++;
                    s.key[freeslot] = v; s.state[freeslot] = 1;
                    return;
                }
                s.fill++; s.used++;
                s.key[e] = v; s.state[e] = 1;
                if (static_cast<unsigned long long>(s.fill) * 5 < mask * 3) return;
                pyset_resize(s, s.used > 50000 ? s.used * 2 : s.used * 4);
                return;
            }
            if (s.state[e] == 1 && s.key[e] == v) return;            // already present
            if (s.state[e] == 2) freeslot = static_cast<int>(e);
            if (probes-- == 0) break;
            ++e;
        }
        perturb >>= 5;
        i = (i * 5 + 1 + perturb) & mask;
    }
}

SS_HD inline bool pyset_discard(PySet& s, long long v) {
    const long long h = pyset_hash(v);
    const unsigned long long mask = static_cast<unsigned long long>(s.mask);
    unsigned long long perturb = static_cast<unsigned long long>(h);
    unsigned long long i = static_cast<unsigned long long>(h) & mask;
    for (;;) {
        unsigned long long e = i;
        int probes = (i + 9 <= mask) ? 9 : 0;
        for (;;) {
            if (s.state[e] == 0) return false;
            if (s.state[e] == 1 && s.key[e] == v) { s.state[e] = 2; s.used--; return true; }
            if (probes-- == 0) break;
            ++e;
        }
        perturb >>= 5;
        i = (i * 5 + 1 + perturb) & mask;
    }
}

// out[] = list(set(values) - {-1}) for ascending `values` (which may contain -1 as first element).  Returns the count,
// or -1 when the emulated table would exceed kPySetMaxTable.
SS_HD inline int pyset_order(const long long* values, int n, long long* out) {
    PySet a;
    pyset_init(a);
    for (int i = 0; i < n; ++i) {
        pyset_add(a, values[i]);
        if (a.overflow) return -1;
    }
    int count = 0;
    if ((a.used >> 2) > 1) {
        // set_copy_and_difference: set_copy = set_merge into an empty set, then discard
        PySet b;
        pyset_init(b);
        if ((b.fill + a.used) * 5 >= b.mask * 3) {
            // set_table_resize(so, (so->used + other->used) * 2) on the empty set
            int newsize = 8;
            while (newsize <= a.used * 2) newsize <<= 1;
            if (newsize > kPySetMaxTable) return -1;
            for (int i = 0; i < newsize; ++i) b.state[i] = 0;
            b.mask = newsize - 1;
        }
        if (b.mask == a.mask && a.fill == a.used) {
            for (int i = 0; i <= a.mask; ++i) { b.key[i] = a.key[i]; b.state[i] = a.state[i]; }
        } else {
            for (int i = 0; i <= a.mask; ++i)
                if (a.state[i] == 1) pyset_insert_clean(b.key, b.state, b.mask, a.key[i]);
        }
        b.fill = a.used; b.used = a.used;
        pyset_discard(b, -1);
        for (int i = 0; i <= b.mask; ++i)
            if (b.state[i] == 1) out[count++] = b.key[i];
        return count;
    }
    PySet r;
    pyset_init(r);
    for (int i = 0; i <= a.mask; ++i)
        if (a.state[i] == 1 && a.key[i] != -1) {
            pyset_add(r, a.key[i]);
            if (r.overflow) return -1;
        }
    for (int i = 0; i <= r.mask; ++i)
        if (r.state[i] == 1) out[count++] = r.key[i];
    return count;
}

// ---------------------------------------------------------------------------------------------------------------
// rectangular linear sum assignment (minimisation), nr x nc costs in row-major doubles
// ---------------------------------------------------------------------------------------------------------------
struct LsapScratch {
    double u[kAssocMaxSide], v[kAssocMaxSide], spc[kAssocMaxSide];
    int path[kAssocMaxSide], col4row[kAssocMaxSide], row4col[kAssocMaxSide], remaining[kAssocMaxSide];
    unsigned char SR[kAssocMaxSide], SC[kAssocMaxSide];
    double tcost[kAssocMaxSide * kAssocMaxSide];
};

// rows_out / cols_out: min(nr, nc) pairs with rows ascending (what linear_sum_assignment returns).  Returns the number
// of pairs, or -1 if infeasible / too large.
SS_HD inline int lsap_solve(const double* cost_in, int nr_in, int nc_in, int* rows_out, int* cols_out, LsapScratch& w) {
    if (nr_in == 0 || nc_in == 0) return 0;
    if (nr_in > kAssocMaxSide || nc_in > kAssocMaxSide) return -1;
    const bool transpose = nc_in < nr_in;
    int nr = nr_in, nc = nc_in;
    const double* cost = cost_in;
    if (transpose) {
        for (int i = 0; i < nr_in; ++i)
            for (int j = 0; j < nc_in; ++j) w.tcost[j * nr_in + i] = cost_in[i * nc_in + j];
        cost = w.tcost;
        nr = nc_in; nc = nr_in;
    }
    const double inf = 1.0 / 0.0;
    for (int i = 0; i < nr; ++i) { w.u[i] = 0.0; w.col4row[i] = -1; }
    for (int j = 0; j < nc; ++j) { w.v[j] = 0.0; w.row4col[j] = -1; w.path[j] = -1; }
    for (int cur = 0; cur < nr; ++cur) {
        // ---- augmenting_path ----
        double min_val = 0.0;
        int i = cur;
        int num_remaining = nc;
        for (int it = 0; it < nc; ++it) w.remaining[it] = nc - it - 1;
        for (int k = 0; k < nr; ++k) w.SR[k] = 0;
        for (int k = 0; k < nc; ++k) { w.SC[k] = 0; w.spc[k] = inf; }
        int sink = -1;
        while (sink == -1) {
            int index = -1;
            double lowest = inf;
            w.SR[i] = 1;
            for (int it = 0; it < num_remaining; ++it) {
                const int j = w.remaining[it];
                const double r = min_val + cost[i * nc + j] - w.u[i] - w.v[j];
                if (r < w.spc[j]) { w.path[j] = i; w.spc[j] = r; }
                if (w.spc[j] < lowest || (w.spc[j] == lowest && w.row4col[j] == -1)) { lowest = w.spc[j]; index = it; }
            }
            min_val = lowest;
            if (min_val == inf) return -1;
            const int j = w.remaining[index];
            if (w.row4col[j] == -1) sink = j;
            else i = w.row4col[j];
            w.SC[j] = 1;
            w.remaining[index] = w.remaining[--num_remaining];
        }
        // ---- dual update ----
        w.u[cur] += min_val;
        for (int k = 0; k < nr; ++k)
            if (w.SR[k] && k != cur) w.u[k] += min_val - w.spc[w.col4row[k]];
        for (int k = 0; k < nc; ++k)
            if (w.SC[k]) w.v[k] -= min_val - w.spc[k];
        // ---- augment ----
        int j = sink;
        for (;;) {
            const int r = w.path[j];
            w.row4col[j] = r;
            const int tmp = w.col4row[r];
            w.col4row[r] = j;
            j = tmp;
            if (r == cur) break;
        }
    }
    int n = 0;
    if (transpose) {
        // pairs (col4row[v], v) sorted by col4row[v] (argsort of a permutation prefix: values are distinct)
        for (int want = 0; want < nr_in; ++want)
            for (int v2 = 0; v2 < nr; ++v2)
                if (w.col4row[v2] == want) { rows_out[n] = want; cols_out[n] = v2; ++n; }
    } else {
        for (int r = 0; r < nr; ++r) { rows_out[n] = r; cols_out[n] = w.col4row[r]; ++n; }
    }
    return n;
}

}  // namespace stemseg
