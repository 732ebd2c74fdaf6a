// Compact KKT record of the quadruped NMPC: only the STRUCTURALLY non-zero slots of the per-node blocks, one contiguous
// 16-byte aligned chunk per shooting node, so that a producer needs one bulk (TMA) store per node and a consumer one or two
// bulk loads per stage.  The dense record (ungar_b200_kkt_layout) spends ~65 % of its bytes on zeros:
//   A_k  13 x 37 = 481 slots, 250 structurally non-zero     (quadruped.example.cpp:162-200: p+/v+ rows see only p_c, v_c and
//                                                            the f_c of each leg; q+/w+ rows see q, w and all 24 inputs)
//   H_k  703 (packed 37 x 37), 61 non-zero                   (:216-244 sums of squares + the per-leg barrier terms of :330-333:
//                                                            13 state diagonals + eight 3 x 3 blocks)
//   C_k  4 x 4 x 20 = 320, 224 non-zero                      (:269-303: a contact row sees p_c, q, r_leg of nodes k and k - 1)
// Chunk layout (doubles; every sub-block starts on an even offset; pads are written as 0):
//   Cs [4 legs][4 rows][8]   columns (p_c, q0..q3, r0..r2) of node k; p_c = p_z for row 0, p_{row-1} for rows 1..3      128
//   Cp [4 legs][3 rows][8]   the same columns of node k - 1, rows 1..3 (row 0 does not see node k - 1)                      96
//   g  [29]  13 defects x_{k+1} - f(x_k, u_k) | 16 contact values                                            (+1 pad)       30
//   q  [37]  QP vector entries of [x_k; u_k]                                                                 (+1 pad)       38
//   Hd [13]  state diagonal of H_k (+1 pad) | Hb [8 blocks][6] upper triangles of the 3 x 3 input blocks                    62
//   h  [12]  inequality values                                                                                              12
//   AQ [7 rows: q+ 0..3, w+ 0..2][32]  columns (q0..q3, w0..w2, pad, u0..u23): the inputs start on an even column       224
//   AP [6 rows: p+ 0..2, v+ 0..2][6]   p+ row c: (A[c][c], A[c][7+c], f_{leg,c} x 4); v+ row c: (A[7+c][7+c], f_{leg,c} x 4, pad)  36
// = 626 doubles (5008 B) per node.  After the N chunks: tail = g0 [13] (+1) | q_N [13] (+1) | HN diagonal [13] (+1) | cost [2].
// Twelve AQ slots per node (d w+_c / d r_{leg,c}) are structurally zero in the reference's CppAD pattern; they are kept so that the
// seven q+/w+ rows share one 32-wide shape.
#pragma once

#include <cstdint>
#include <vector>

namespace ub {

struct Compact {
    static constexpr int oCs = 0, oCp = 128, oG = 224, oQ = 254, oHd = 292, oHb = 306, oHi = 354, oAQ = 366, oAP = 590, NODE = 626;
    static constexpr int SMALL = oHi;        // what a QP stage needs of the first part (Cs, Cp, g, q, H): 354 doubles = 2832 B
    static constexpr int APART = NODE - oAQ;  // AQ + AP: 260 doubles = 2080 B
    static constexpr int tG0 = 0, tQN = 14, tHN = 28, tCost = 42, TAIL = 44;
    static_assert((NODE * 8) % 16 == 0 && (SMALL * 8) % 16 == 0 && (oAQ * 8) % 16 == 0 && (APART * 8) % 16 == 0, "bulk copies need 16-byte multiples");
    __host__ __device__ static constexpr long long size(int N) { return (long long)N * NODE + TAIL; }
    __host__ __device__ static constexpr long long tail(int N) { return (long long)N * NODE; }
    // AQ column of local variable z (0..36), or -1 when the column lies outside the q+/w+ pattern (p and v columns)
    __host__ __device__ static constexpr int aq_col(int z) { return z < 3 ? -1 : z < 7 ? z - 3 : z < 10 ? -1 : z < 13 ? z - 6 : z - 5; }
    __host__ __device__ static constexpr int aq_row(int row) { return row >= 3 && row < 7 ? row - 3 : row >= 10 ? row - 6 : -1; }
};

// Host-side: for every slot of the compact record its offset in the dense record (ungar_b200_kkt_layout), or -2 for a pad
// (the gather convention of sweep.cuh::gather_kernel: -2 reads as 0).  Dense offsets follow make_layout / DESIGN.md §3.
struct DenseOffsets {
    long long g, A, C, h, cost, grad, H, HN;
};

inline int tri37(int i, int j) { return i * 37 - (i * (i - 1)) / 2 + (j - i); }

inline std::vector<int32_t> compact_to_dense_map(int N, const DenseOffsets& D) {
    using K = Compact;
    std::vector<int32_t> map(size_t(K::size(N)), -2);
    const int nX = 13 * (N + 1);
    for (int k = 0; k < N; ++k) {
        int32_t* m = map.data() + size_t(k) * K::NODE;
        for (int leg = 0; leg < 4; ++leg)
            for (int rr = 0; rr < 4; ++rr) {
                const long long row = D.C + ((long long)(k * 4 + leg) * 4 + rr) * 20;
                const int pc = rr == 0 ? 2 : rr - 1;
                int32_t* cs = m + K::oCs + (leg * 4 + rr) * 8;
                cs[0] = int32_t(row + pc);
                for (int i = 0; i < 4; ++i) cs[1 + i] = int32_t(row + 3 + i);
                for (int i = 0; i < 3; ++i) cs[5 + i] = int32_t(row + 7 + i);
                if (rr > 0) {
                    int32_t* cp = m + K::oCp + (leg * 3 + rr - 1) * 8;
                    cp[0] = int32_t(row + 10 + pc);
                    for (int i = 0; i < 4; ++i) cp[1 + i] = int32_t(row + 13 + i);
                    for (int i = 0; i < 3; ++i) cp[5 + i] = int32_t(row + 17 + i);
                }
            }
        for (int i = 0; i < 13; ++i) m[K::oG + i] = int32_t(D.g + 13 + 13 * k + i);
        for (int i = 0; i < 16; ++i) m[K::oG + 13 + i] = int32_t(D.g + nX + 16 * k + i);
        for (int i = 0; i < 13; ++i) m[K::oQ + i] = int32_t(D.grad + 13 * k + i);
        for (int i = 0; i < 24; ++i) m[K::oQ + 13 + i] = int32_t(D.grad + nX + 24 * k + i);
        const long long Hk = D.H + (long long)k * 703;
        for (int i = 0; i < 13; ++i) m[K::oHd + i] = int32_t(Hk + tri37(i, i));
        for (int b = 0; b < 8; ++b) {
            const int a = 13 + 3 * b;
            int32_t* hb = m + K::oHb + 6 * b;
            hb[0] = int32_t(Hk + tri37(a, a)); hb[1] = int32_t(Hk + tri37(a, a + 1)); hb[2] = int32_t(Hk + tri37(a, a + 2));
            hb[3] = int32_t(Hk + tri37(a + 1, a + 1)); hb[4] = int32_t(Hk + tri37(a + 1, a + 2)); hb[5] = int32_t(Hk + tri37(a + 2, a + 2));
        }
        for (int i = 0; i < 12; ++i) m[K::oHi + i] = int32_t(D.h + 12 * k + i);
        const long long Ak = D.A + (long long)k * 481;
        for (int row = 0; row < 13; ++row) {
            const int qr = K::aq_row(row);
            if (qr >= 0) {
                for (int z = 0; z < 37; ++z)
                    if (K::aq_col(z) >= 0) m[K::oAQ + qr * 32 + K::aq_col(z)] = int32_t(Ak + row * 37 + z);
            } else if (row < 3) {  // p+ row c
                int32_t* ap = m + K::oAP + row * 6;
                ap[0] = int32_t(Ak + row * 37 + row); ap[1] = int32_t(Ak + row * 37 + 7 + row);
                for (int leg = 0; leg < 4; ++leg) ap[2 + leg] = int32_t(Ak + row * 37 + 13 + 6 * leg + row);
            } else {  // v+ row c = row - 7
                const int c = row - 7;
                int32_t* ap = m + K::oAP + (3 + c) * 6;
                ap[0] = int32_t(Ak + row * 37 + row);
                for (int leg = 0; leg < 4; ++leg) ap[1 + leg] = int32_t(Ak + row * 37 + 13 + 6 * leg + c);
            }
        }
    }
    int32_t* t = map.data() + K::tail(N);
    for (int i = 0; i < 13; ++i) {
        t[K::tG0 + i] = int32_t(D.g + i);
        t[K::tQN + i] = int32_t(D.grad + 13 * N + i);
        t[K::tHN + i] = int32_t(D.HN + (i * 13 - (i * (i - 1)) / 2));
    }
    t[K::tCost] = int32_t(D.cost); t[K::tCost + 1] = int32_t(D.cost + 1);
    return map;
}

}  // namespace ub
