// tq_kernels.cu -- see tq_kernels.cuh.  Compiled only for sm_100a.
#include "tq_kernels.cuh"

namespace tq {
namespace {

#ifndef TQ_MIN_BLOCKS
#define TQ_MIN_BLOCKS 2
#endif
constexpr int NA = 1 << kRegBits;  // amplitudes per thread
constexpr int kOpSlots = 48;       // ops (+ their 256-byte matrix / term slots) resident in shared memory
constexpr int kWinSlots = 32;      // window headers resident in shared memory
static_assert(kOpSlots >= kMaxWindowOps, "a window's ops must fit the staged range");

// ---- bank swizzle: slot(j) = j ^ fold(j >> 3); fold is GF(2)-linear, so slot(a ^ b) = slot(a) ^ slot(b) ----------
constexpr uint32_t swz_mask(int out_bit) {
    uint32_t m = 0;
    for (int p = 3; p < 13; ++p)
        if ((kSwizzleVec[p] >> out_bit) & 1) m |= 1u << (p - 3);
    return m;
}
constexpr uint32_t kSwzM0 = swz_mask(0), kSwzM1 = swz_mask(1), kSwzM2 = swz_mask(2);

__device__ __forceinline__ uint32_t swz(uint32_t j) {
    const uint32_t h = j >> 3;
    return j ^ ((__popc(h & kSwzM0) & 1u) | ((__popc(h & kSwzM1) & 1u) << 1) | ((__popc(h & kSwzM2) & 1u) << 2));
}

typedef double2 Amps[NA];

// ---- register-window blocks.  RB.. = register bits; all loops are fully unrolled (static register indices) -------
__device__ __forceinline__ double2 cmul(double2 m, double2 v) {
    return make_double2(m.x * v.x - m.y * v.y, m.x * v.y + m.y * v.x);
}
__device__ __forceinline__ double2 cfma(double2 m, double2 v, double2 acc) {
    return make_double2(fma(m.x, v.x, fma(-m.y, v.y, acc.x)), fma(m.x, v.y, fma(m.y, v.x, acc.y)));
}

// dense 4x4 block on register bits RA < RB (matrix index bit 0 = RA); m = 16 staged entries, row-major.
// Two groups of four amplitudes per sweep over the matrix, column by column: each column's four entries update all
// eight outputs, i.e. sixteen independent FMA chains in flight; inputs are read in place, outputs collected in
// temporaries and written back at the end of the sweep.
template <int RA, int RB>
__device__ __forceinline__ void g_u2(Amps& a, const double2* __restrict__ m) {
    constexpr int O0 = (RA != 0 && RB != 0) ? 0 : (RA != 1 && RB != 1) ? 1 : 2;
    constexpr int O1 = (RA != 3 && RB != 3) ? 3 : (RA != 2 && RB != 2) ? 2 : 1;
#pragma unroll
    for (int gp = 0; gp < 2; ++gp) {
        const int base0 = (gp ? (1 << O1) : 0), base1 = base0 | (1 << O0);
        double2 o0[4], o1[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int off = ((k & 1) ? (1 << RA) : 0) | ((k & 2) ? (1 << RB) : 0);
            const double2 x0 = a[base0 | off], x1 = a[base1 | off];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double2 mm = m[i * 4 + k];
                if (k == 0) {
                    o0[i] = make_double2(mm.x * x0.x, mm.x * x0.y);
                    o1[i] = make_double2(mm.x * x1.x, mm.x * x1.y);
                } else {
                    o0[i].x = fma(mm.x, x0.x, o0[i].x);
                    o0[i].y = fma(mm.x, x0.y, o0[i].y);
                    o1[i].x = fma(mm.x, x1.x, o1[i].x);
                    o1[i].y = fma(mm.x, x1.y, o1[i].y);
                }
                o0[i].x = fma(-mm.y, x0.y, o0[i].x);
                o0[i].y = fma(mm.y, x0.x, o0[i].y);
                o1[i].x = fma(-mm.y, x1.y, o1[i].x);
                o1[i].y = fma(mm.y, x1.x, o1[i].y);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int off = ((i & 1) ? (1 << RA) : 0) | ((i & 2) ? (1 << RB) : 0);
            a[base0 | off] = o0[i];
            a[base1 | off] = o1[i];
        }
    }
}
// dense 2x2 block on register bit RB; m = {m00, m01, m10, m11}
template <int RB>
__device__ __forceinline__ void g_u1(Amps& a, const double2* __restrict__ m) {
    const double2 m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma unroll
    for (int p = 0; p < NA / 2; ++p) {
        const int i0 = ((p >> RB) << (RB + 1)) | (p & ((1 << RB) - 1)), i1 = i0 | (1 << RB);
        const double2 a0 = a[i0], a1 = a[i1];
        a[i0] = cfma(m01, a1, cmul(m00, a0));
        a[i1] = cfma(m11, a1, cmul(m10, a0));
    }
}
// diagonal block on register bit RB: m[0] on bit = 0, m[3] on bit = 1
template <int RB>
__device__ __forceinline__ void g_d1(Amps& a, const double2* __restrict__ m) {
    const double2 d0 = m[0], d1 = m[3];
#pragma unroll
    for (int i = 0; i < NA; ++i) a[i] = cmul(((i >> RB) & 1) ? d1 : d0, a[i]);
}
__device__ __forceinline__ void g_scale(Amps& a, double2 d) {
#pragma unroll
    for (int i = 0; i < NA; ++i) a[i] = cmul(d, a[i]);
}
template <int RB>
__device__ __forceinline__ void g_x_if(Amps& a, bool pred) {  // X on RB where pred (per-thread control bit)
#pragma unroll
    for (int p = 0; p < NA / 2; ++p) {
        const int i0 = ((p >> RB) << (RB + 1)) | (p & ((1 << RB) - 1)), i1 = i0 | (1 << RB);
        const double2 a0 = a[i0], a1 = a[i1];
        a[i0] = pred ? a1 : a0;
        a[i1] = pred ? a0 : a1;
    }
}
template <int CB, int TB>
__device__ __forceinline__ void g_cx(Amps& a) {  // both bits in the window: a register permutation
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        if (((i >> CB) & 1) && !((i >> TB) & 1)) {
            const double2 t = a[i];
            a[i] = a[i | (1 << TB)];
            a[i | (1 << TB)] = t;
        }
    }
}
// exact 1-qubit depolarising channel on (row bit A, column bit B): rho -> (1-4p/3) rho + (2p/3) Tr_q(rho) (x) I
template <int A, int B>
__device__ __forceinline__ void g_depol1(Amps& a, double p) {
    const double keep = 1.0 - 2.0 * p / 3.0, mixw = 2.0 * p / 3.0, off = 1.0 - 4.0 * p / 3.0;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        if (((i >> A) & 1) || ((i >> B) & 1)) continue;
        const int i11 = i | (1 << A) | (1 << B), i10 = i | (1 << A), i01 = i | (1 << B);
        const double2 r00 = a[i], r11 = a[i11];
        a[i] = make_double2(keep * r00.x + mixw * r11.x, keep * r00.y + mixw * r11.y);
        a[i11] = make_double2(keep * r11.x + mixw * r00.x, keep * r11.y + mixw * r00.y);
        a[i10] = make_double2(off * a[i10].x, off * a[i10].y);
        a[i01] = make_double2(off * a[i01].x, off * a[i01].y);
    }
}
// exact 2-qubit depolarising channel; window = {P0, P1} (row/col of one qubit) + {P2, P3} (row/col of the other):
// rho -> (1 - 16p/15) rho + (4p/15) Tr_ab(rho) (x) I_4
template <int P0, int P1, int P2, int P3>
__device__ __forceinline__ void g_depol2(Amps& a, double p) {
    const double alpha = 1.0 - 16.0 * p / 15.0, beta = 4.0 * p / 15.0;
    double2 tr = make_double2(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < NA; ++i)
        if (((i >> P0) & 1) == ((i >> P1) & 1) && ((i >> P2) & 1) == ((i >> P3) & 1)) { tr.x += a[i].x; tr.y += a[i].y; }
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        double2 v = make_double2(alpha * a[i].x, alpha * a[i].y);
        if (((i >> P0) & 1) == ((i >> P1) & 1) && ((i >> P2) & 1) == ((i >> P3) & 1)) { v.x += beta * tr.x; v.y += beta * tr.y; }
        a[i] = v;
    }
}

#define TQ_RB4(FN, rb, ...)                                        \
    switch (rb) {                                                  \
    case 0: FN<0>(__VA_ARGS__); break;                             \
    case 1: FN<1>(__VA_ARGS__); break;                             \
    case 2: FN<2>(__VA_ARGS__); break;                             \
    default: FN<3>(__VA_ARGS__); break;                            \
    }

__device__ __forceinline__ void exec_cx_ww(Amps& a, int cb, int tb) {
    switch (cb * 4 + tb) {
    case 1: g_cx<0, 1>(a); break;
    case 2: g_cx<0, 2>(a); break;
    case 3: g_cx<0, 3>(a); break;
    case 4: g_cx<1, 0>(a); break;
    case 6: g_cx<1, 2>(a); break;
    case 7: g_cx<1, 3>(a); break;
    case 8: g_cx<2, 0>(a); break;
    case 9: g_cx<2, 1>(a); break;
    case 11: g_cx<2, 3>(a); break;
    case 12: g_cx<3, 0>(a); break;
    case 13: g_cx<3, 1>(a); break;
    case 14: g_cx<3, 2>(a); break;
    default: break;
    }
}

__device__ __forceinline__ void exec_u2(Amps& a, int ra, int rb, const double2* m) {
    switch (ra * 4 + rb) {
    case 1: g_u2<0, 1>(a, m); break;
    case 2: g_u2<0, 2>(a, m); break;
    case 3: g_u2<0, 3>(a, m); break;
    case 6: g_u2<1, 2>(a, m); break;
    case 7: g_u2<1, 3>(a, m); break;
    default: g_u2<2, 3>(a, m); break;
    }
}

__device__ __forceinline__ void exec_depol1(Amps& a, int ra, int rb, double p) {
    const int lo = ra < rb ? ra : rb, hi = ra < rb ? rb : ra;
    switch (lo * 4 + hi) {
    case 1: g_depol1<0, 1>(a, p); break;
    case 2: g_depol1<0, 2>(a, p); break;
    case 3: g_depol1<0, 3>(a, p); break;
    case 6: g_depol1<1, 2>(a, p); break;
    case 7: g_depol1<1, 3>(a, p); break;
    default: g_depol1<2, 3>(a, p); break;
    }
}

__device__ __forceinline__ void exec_depol2(Amps& a, int rows, int cols, double p) {
    // qubit a: row bit rows & 3, column bit cols & 3; the partner of register bit 0 decides the pairing
    const int ra = rows & 3, ca = cols & 3, rb = (rows >> 2) & 3, cb = (cols >> 2) & 3;
    const int partner0 = (ra == 0) ? ca : (ca == 0) ? ra : (rb == 0) ? cb : rb;
    if (partner0 == 1) g_depol2<0, 1, 2, 3>(a, p);
    else if (partner0 == 2) g_depol2<0, 2, 1, 3>(a, p);
    else g_depol2<0, 3, 1, 2>(a, p);
}

// ---- expectation on registers (see W_EXPC / W_EXPD in tq_plan.h) ------------------------------------------------
// one class of an off-diagonal group: XR = flip mask over the register bits; d[0] = Z/Y mask outside the window,
// d[1..8] = (cA, cB) per register pair
template <int XR, bool IMAG>
__device__ __forceinline__ double g_expc(const Amps& a, uint64_t ctx, const double2* __restrict__ d) {
    double sum = 0.0;
    int q = 0;
#pragma unroll
    for (int r = 0; r < NA; ++r) {
        if ((r ^ XR) > r) {
            const double2 v = a[r], w = a[r ^ XR], c = d[1 + q];
            sum = fma(c.x, w.x * v.x + w.y * v.y, sum);                 // cA * Re(conj(psi[r^x]) psi[r])
            if (IMAG) sum = fma(-c.y, w.x * v.y - w.y * v.x, sum);      // -cB * Im(...)
            ++q;
        }
    }
    const uint64_t zphys = (uint64_t)__double_as_longlong(d[0].x);
    return (__popcll(ctx & zphys) & 1) ? -sum : sum;
}

// all diagonal terms: Walsh-Hadamard transform of the 16 probabilities, then one signed weight sum per class
__device__ __forceinline__ double g_expd(const Amps& a, uint64_t ctx, const double2* __restrict__ head,
                                         const double2* __restrict__ terms) {
    double n[NA];
#pragma unroll
    for (int r = 0; r < NA; ++r) n[r] = a[r].x * a[r].x + a[r].y * a[r].y;
#pragma unroll
    for (int bitp = 0; bitp < kRegBits; ++bitp)
#pragma unroll
        for (int r = 0; r < NA; ++r)
            if (!((r >> bitp) & 1)) {
                const double x = n[r], y = n[r | (1 << bitp)];
                n[r] = x + y;
                n[r | (1 << bitp)] = x - y;
            }
    const unsigned short* cnt = reinterpret_cast<const unsigned short*>(head);
    double total = 0.0;
    int idx = 0;
#pragma unroll
    for (int zr = 0; zr < NA; ++zr) {
        const int c = cnt[zr];
        double s = 0.0;
        for (int i = 0; i < c; ++i) {
            const double2 t = __ldg(terms + idx + i);
            s += (__popcll(ctx & (uint64_t)__double_as_longlong(t.x)) & 1) ? -t.y : t.y;
        }
        idx += c;
        total = fma(s, n[zr], total);
    }
    return total;
}

template <bool IMAG>
__device__ __forceinline__ double exec_expc(const Amps& a, uint64_t ctx, int xr, const double2* d) {
    switch (xr) {
    case 1: return g_expc<1, IMAG>(a, ctx, d);
    case 2: return g_expc<2, IMAG>(a, ctx, d);
    case 3: return g_expc<3, IMAG>(a, ctx, d);
    case 4: return g_expc<4, IMAG>(a, ctx, d);
    case 5: return g_expc<5, IMAG>(a, ctx, d);
    case 6: return g_expc<6, IMAG>(a, ctx, d);
    case 7: return g_expc<7, IMAG>(a, ctx, d);
    case 8: return g_expc<8, IMAG>(a, ctx, d);
    case 9: return g_expc<9, IMAG>(a, ctx, d);
    case 10: return g_expc<10, IMAG>(a, ctx, d);
    case 11: return g_expc<11, IMAG>(a, ctx, d);
    case 12: return g_expc<12, IMAG>(a, ctx, d);
    case 13: return g_expc<13, IMAG>(a, ctx, d);
    case 14: return g_expc<14, IMAG>(a, ctx, d);
    default: return g_expc<15, IMAG>(a, ctx, d);
    }
}

// deterministic CTA-wide sum (fixed shuffle tree, then warps added in index order); result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* s_red, int tid, int nthreads) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
    __syncthreads();
    double total = 0.0;
    if (tid == 0) {
        const int nw = (nthreads + 31) >> 5;
        for (int w = 0; w < nw; ++w) total += s_red[w];
    }
    return total;
}

// ---- block matrices: the matrix program of one fused block for one batch element -------------------------------
__device__ __forceinline__ void eval_block_matrix(const MatDesc md, const MatGate* __restrict__ prog,
                                                  const double* __restrict__ params, size_t b, int ld_params,
                                                  const uint8_t* __restrict__ codes, int ld_codes,
                                                  double2* __restrict__ out) {
    double2 M[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) M[r][c] = make_double2(r == c ? 1.0 : 0.0, 0.0);
    for (int gi = md.begin; gi < md.end; ++gi) {
        const MatGate g = prog[gi];
        int kind = g.kind;
        if (kind == MG_CX) {  // control = lq: swap the two rows with the control bit set
            if (g.lq == 0) {
#pragma unroll
                for (int c = 0; c < 4; ++c) { const double2 t = M[1][c]; M[1][c] = M[3][c]; M[3][c] = t; }
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) { const double2 t = M[2][c]; M[2][c] = M[3][c]; M[3][c] = t; }
            }
            continue;
        }
        if (kind == MG_PAULI_SLOT) {
            const int code = (codes[(size_t)b * ld_codes + g.pidx] >> (int)g.fixed) & 3;
            if (code == 0) continue;
            kind = MG_X + code - 1;
        }
        double2 g00 = make_double2(0.0, 0.0), g01 = g00, g10 = g00, g11 = g00;
        if (kind <= MG_RZ) {
            const double theta = g.pidx >= 0 ? params[(size_t)b * ld_params + g.pidx] : g.fixed;
            double s, c;
            sincos(0.5 * theta, &s, &c);
            if (kind == MG_RX) { g00.x = c; g01.y = s; g10.y = s; g11.x = c; }            // cos I + i sin X
            else if (kind == MG_RY) { g00.x = c; g01.x = s; g10.x = -s; g11.x = c; }      // cos I + i sin Y
            else { g00.x = c; g00.y = s; g11.x = c; g11.y = -s; }                         // diag(e^{+it/2}, e^{-it/2})
        } else if (kind == MG_X) { g01.x = 1.0; g10.x = 1.0; }
        else if (kind == MG_Y) { g01.y = -1.0; g10.y = 1.0; }
        else { g00.x = 1.0; g11.x = -1.0; }
        // M <- (G on block qubit lq) * M : mixes row pairs that differ in bit lq
        if (g.lq == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int r = 0; r < 4; r += 2) {
                    const double2 x0 = M[r][c], x1 = M[r + 1][c];
                    M[r][c] = cfma(g01, x1, cmul(g00, x0));
                    M[r + 1][c] = cfma(g11, x1, cmul(g10, x0));
                }
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const double2 x0 = M[r][c], x1 = M[r + 2][c];
                    M[r][c] = cfma(g01, x1, cmul(g00, x0));
                    M[r + 2][c] = cfma(g11, x1, cmul(g10, x0));
                }
        }
    }
    if (md.nq == 2) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) out[r * 4 + c] = M[r][c];
    } else {  // one-qubit block: the 2x2 matrix in entries 0..3
        out[0] = M[0][0];
        out[1] = M[0][1];
        out[2] = M[1][0];
        out[3] = M[1][1];
    }
}

// One COLUMN of a block matrix (a gate mixes rows, so the four columns never meet and each sees exactly the arithmetic of
// eval_block_matrix: same results, bit for bit), with the program and the sin / cos of its rotations already in shared memory.
__device__ __forceinline__ void eval_block_column_staged(const MatDesc md, const MatGate* s_prog, const double2* s_sc,
                                                         const uint8_t* __restrict__ codes, int col,
                                                         double2* __restrict__ out) {
    double2 M[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) M[r] = make_double2(r == col ? 1.0 : 0.0, 0.0);
    for (int gi = md.begin; gi < md.end; ++gi) {
        const MatGate g = s_prog[gi];
        int kind = g.kind;
        if (kind == MG_CX) {
            const int ra = g.lq == 0 ? 1 : 2;
            const double2 t = M[ra]; M[ra] = M[3]; M[3] = t;
            continue;
        }
        if (kind == MG_PAULI_SLOT) {
            const int code = (codes[g.pidx] >> (int)g.fixed) & 3;
            if (code == 0) continue;
            kind = MG_X + code - 1;
        }
        double2 g00 = make_double2(0.0, 0.0), g01 = g00, g10 = g00, g11 = g00;
        if (kind <= MG_RZ) {
            const double sn = s_sc[gi].x, cs = s_sc[gi].y;
            if (kind == MG_RX) { g00.x = cs; g01.y = sn; g10.y = sn; g11.x = cs; }
            else if (kind == MG_RY) { g00.x = cs; g01.x = sn; g10.x = -sn; g11.x = cs; }
            else { g00.x = cs; g00.y = sn; g11.x = cs; g11.y = -sn; }
        } else if (kind == MG_X) { g01.x = 1.0; g10.x = 1.0; }
        else if (kind == MG_Y) { g01.y = -1.0; g10.y = 1.0; }
        else { g00.x = 1.0; g11.x = -1.0; }
        if (g.lq == 0) {
#pragma unroll
            for (int r = 0; r < 4; r += 2) {
                const double2 x0 = M[r], x1 = M[r + 1];
                M[r] = cfma(g01, x1, cmul(g00, x0));
                M[r + 1] = cfma(g11, x1, cmul(g10, x0));
            }
        } else {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const double2 x0 = M[r], x1 = M[r + 2];
                M[r] = cfma(g01, x1, cmul(g00, x0));
                M[r + 2] = cfma(g11, x1, cmul(g10, x0));
            }
        }
    }
    if (md.nq == 2) {
#pragma unroll
        for (int r = 0; r < 4; ++r) out[r * 4 + col] = M[r];
    } else if (col < 2) {  // one-qubit block: the 2x2 matrix in entries 0..3
        out[col] = M[0];
        out[2 + col] = M[1];
    }
}

// Single-tile plans fold prep_matrices_kernel into the pass: the CTA of element b evaluates its block matrices itself.  This
// is the latency path (one COBYLA cost evaluation per call: the reference's own loop), and what it costs is a chain of
// dependent memory accesses, not arithmetic -- so: the element's angles (they may live in pinned HOST memory), the matrix
// program and the block descriptors come in with coalesced sweeps, every rotation's sin / cos is evaluated by its own thread,
// and four threads share a block (one column each).  One thread per block reading the program gate by gate from global
// memory, with a sincos per gate in the chain, was 22.5 / 54 / 38 us of kernel time for the 4 / 6 / 8-qubit bench shapes at
// B = 1 (ncu).  `scratch`: kOpSlots * kMatStride double2 of shared memory that nothing else uses yet.
__device__ __forceinline__ void fused_prep(const PassParams& p, uint32_t b, int tid, int nthreads, double2* scratch,
                                           double2* __restrict__ out_mats) {
    const double* par = p.params ? p.params + (size_t)b * p.ld_params : nullptr;
    const uint8_t* cod = p.codes ? p.codes + (size_t)b * p.ld_codes : nullptr;
    const size_t cap = (size_t)kOpSlots * kMatStride * sizeof(double2);
    const size_t off_prog = (size_t)p.ld_params * sizeof(double);
    const size_t off_sc = (off_prog + (size_t)p.n_prog * sizeof(MatGate) + 15) & ~(size_t)15;
    const size_t need = off_sc + (size_t)p.n_prog * sizeof(double2);
    if (p.n_prog > 0 && need <= cap) {
        unsigned char* raw = reinterpret_cast<unsigned char*>(scratch);
        double* s_par = reinterpret_cast<double*>(raw);
        MatGate* s_prog = reinterpret_cast<MatGate*>(raw + off_prog);
        double2* s_sc = reinterpret_cast<double2*>(raw + off_sc);
        if (par)
            for (int i = tid; i < p.ld_params; i += nthreads) s_par[i] = par[i];
        {   // the program as 8-byte words (24-byte gates)
            const unsigned long long* src = reinterpret_cast<const unsigned long long*>(p.prog);
            unsigned long long* dst = reinterpret_cast<unsigned long long*>(s_prog);
            for (int i = tid; i < p.n_prog * 3; i += nthreads) dst[i] = __ldg(src + i);
        }
        __syncthreads();
        for (int gi = tid; gi < p.n_prog; gi += nthreads) {
            const MatGate g = s_prog[gi];
            if (g.kind <= MG_RZ) {
                const double theta = g.pidx >= 0 ? s_par[g.pidx] : g.fixed;
                double sn, cs;
                sincos(0.5 * theta, &sn, &cs);
                s_sc[gi] = make_double2(sn, cs);
            }
        }
        __syncthreads();
        for (int idx = tid; idx < 4 * p.n_mats; idx += nthreads)
            eval_block_column_staged(p.descs[idx >> 2], s_prog, s_sc, cod, idx & 3, out_mats + (size_t)(idx >> 2) * kMatStride);
    } else {
        if (par && (size_t)p.ld_params * sizeof(double) <= cap) {
            double* s_par = reinterpret_cast<double*>(scratch);
            for (int i = tid; i < p.ld_params; i += nthreads) s_par[i] = par[i];
            __syncthreads();
            par = s_par;
        }
        for (int mi = tid; mi < p.n_mats; mi += nthreads)
            eval_block_matrix(p.descs[mi], p.prog, par, 0, p.ld_params, cod, p.ld_codes, out_mats + (size_t)mi * kMatStride);
    }
    __syncthreads();
}

#define TQ_SEL4(i, v0, v1, v2, v3, OP) ((((i) & 1) ? (v0) : 0u) OP (((i) & 2) ? (v1) : 0u) OP (((i) & 4) ? (v2) : 0u) OP (((i) & 8) ? (v3) : 0u))
#define TQ_SLOT(r) (slot_t ^ TQ_SEL4(r, ws0, ws1, ws2, ws3, ^))
#define TQ_IO_SLOT(i) (io_slot ^ TQ_SEL4(i, iw0, iw1, iw2, iw3, ^))
#define TQ_IO_GOFF(i) (io_goff | TQ_SEL4(i, ig0, ig1, ig2, ig3, |))

// flat dispatch ids (computed while staging a window's ops; the planner's WinOp codes stay symbolic)
enum : int { F_U2 = 0, F_U1 = 6, F_D1 = 10, F_D1_OUT = 14, F_CX_WW = 15, F_CX_OW = 31, F_DEPOL1 = 35, F_DEPOL2 = 36, F_EXPC = 37, F_EXPC_IMAG = 38, F_EXPD = 39 };

__device__ __forceinline__ int flat_code(uint32_t w0) {
    const int code = w0 & 0xff, rb = (w0 >> 8) & 0xf, rb2 = (w0 >> 12) & 0xf;
    switch (code) {
    case W_U2: return F_U2 + (rb == 0 ? rb2 - 1 : rb == 1 ? rb2 + 1 : 5);  // (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
    case W_U1: return F_U1 + rb;
    case W_D1: return F_D1 + rb;
    case W_D1_OUT: return F_D1_OUT;
    case W_CX_OW: return F_CX_OW + rb;
    case W_DEPOL1: return F_DEPOL1;
    case W_DEPOL2: return F_DEPOL2;
    case W_EXPC: return (rb2 & 1) ? F_EXPC_IMAG : F_EXPC;
    default: return F_EXPD;
    }
}

// TABLE = true: every CTA runs its OWN single-tile problem (circuit, block matrices, Hamiltonian, initial state): CTA i
// takes its parameters from p_in.table[i] -- B different environments evaluated by one launch.
template <bool DM, bool TABLE>
__global__ void __launch_bounds__(kMaxThreads, TQ_MIN_BLOCKS) tile_pass_kernel(const PassParams p_in) {
    PassParams p_own;
    if (TABLE) p_own = p_in.table[blockIdx.x];
    const PassParams& p = TABLE ? p_own : p_in;
    const uint32_t cta = TABLE ? 0u : blockIdx.x;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tile_amps = 1 << p.k_eff;
    const int valid_amps = 1 << p.k;
    double2* amp = reinterpret_cast<double2*>(smem_raw);
    double2* s_mat = amp + tile_amps;                                          // kOpSlots x 16 entries
    WinOp* s_wops = reinterpret_cast<WinOp*>(s_mat + kOpSlots * kMatStride);   // kOpSlots entries
    double* s_red = reinterpret_cast<double*>(s_wops + kOpSlots);              // 32 entries
    uint2* s_win = reinterpret_cast<uint2*>(s_red + 32);                       // kWinSlots x 3 (24-byte headers)
    uint32_t* hi_off = reinterpret_cast<uint32_t*>(s_win + 3 * kWinSlots);     // 2^(k - lead) entries

    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int nar = (p.arith_threads > 0 && p.arith_threads < nthreads) ? p.arith_threads : nthreads;   // (see PassParams)
    const uint32_t ntiles = 1u << p.n_nl;
    const uint32_t tile = cta & (ntiles - 1u);
    const uint32_t b = cta >> p.n_nl;

    uint64_t tile_base = 0;
    for (int i = 0; i < p.n_nl; ++i) tile_base |= (uint64_t)((tile >> i) & 1u) << p.nonlocal[i];

    // hi_off[h]: physical offset of the tile positions above the contiguous lead run
    const int n_hi = 1 << (p.k - p.lead);
    for (int h = tid; h < n_hi; h += nthreads) {
        uint32_t off = 0;
        for (int i = p.lead; i < p.k; ++i) off |= ((uint32_t)(h >> (i - p.lead)) & 1u) << p.local[i];
        hi_off[h] = off;
    }
    __syncthreads();
    const uint32_t lead_mask = (1u << p.lead) - 1u;
    // physical offset of tile index j (j < 2^k)
#define TQ_PHYS(j) ((uint32_t)(((j) & lead_mask) | hi_off[(j) >> p.lead]))

    // I/O layout: amplitude i of thread t is tile index t + i * nthreads; both the swizzled slot and the physical
    // offset are bitwise-linear in the index, so they split into a per-thread part and four per-bit constants
    const uint32_t io_slot = swz(tid);
    const uint32_t iw0 = swz(nthreads), iw1 = swz(nthreads << 1), iw2 = swz(nthreads << 2), iw3 = swz(nthreads << 3);
    const uint32_t io_goff = tid < valid_amps ? TQ_PHYS(tid) : 0u;
    const uint32_t ig0 = nthreads < valid_amps ? TQ_PHYS(nthreads) : 0u;
    const uint32_t ig1 = (nthreads << 1) < valid_amps ? TQ_PHYS(nthreads << 1) : 0u;
    const uint32_t ig2 = (nthreads << 2) < valid_amps ? TQ_PHYS(nthreads << 2) : 0u;
    const uint32_t ig3 = (nthreads << 3) < valid_amps ? TQ_PHYS(nthreads << 3) : 0u;

    const uint64_t elem_off = (uint64_t)b << p.nbits;

    const double2* my_mats = p.mats + (size_t)b * p.n_mats * kMatStride;
    if (p.fused_prep)   // single-tile plans: this CTA is the only one of its element and evaluates its block matrices itself
        fused_prep(p, b, tid, nthreads, s_mat, const_cast<double2*>(my_mats));
    const bool active = tid < (tile_amps >> kRegBits);
    const int n_tbits = p.k_eff - kRegBits;
    const int n_run = (!DM && p.exp_mode == 1) ? p.n_windows : p.n_gate_windows;

    // ops [staged_begin, staged_end) of the pass live in shared memory with their block matrices / term lists
    int staged_begin = 0, staged_end = 0;
    auto stage_ops = [&](int first) {
        staged_begin = first;
        staged_end = min(first + kOpSlots, p.n_wops);
        for (int i = tid; i < (staged_end - staged_begin) * kMatStride; i += nthreads) {
            const int oi = i >> 4, e = i & 15;
            WinOp wo = p.wops[staged_begin + oi];
            const int code = wo.w0 & 0xff, flags = wo.w0 >> 24;
            if (e == 0) {
                wo.w0 = (wo.w0 & ~0xffu) | (uint32_t)flat_code(wo.w0);
                s_wops[oi] = wo;
            }
            if (code >= W_EXPC) {  // expectation data (qsel 16-byte units) rides in the op's matrix slot
                if (e < (int)((wo.w0 >> 16) & 0xff))
                    s_mat[oi * kMatStride + e] = reinterpret_cast<const double2*>(p.eterms)[(size_t)wo.t + e];
            } else if (code <= W_D1_OUT && (code == W_U2 || e < 4)) {
                int src = e;
                if (code == W_U2 && (flags & FLAG_SWAP)) {  // matrix written for the other qubit order
                    const int r = e >> 2, c = e & 3;
                    src = ((((r & 1) << 1) | (r >> 1)) << 2) | ((c & 1) << 1) | (c >> 1);
                }
                double2 v = my_mats[(size_t)wo.t * kMatStride + src];
                if (flags & FLAG_CONJ) v.y = -v.y;
                s_mat[oi * kMatStride + e] = v;
            }
        }
    };

    // ---- 1. stage the tile (coalesced global reads, swizzled shared-memory slots), the window headers and the
    //         first kOpSlots ops of the pass; all global loads are in flight together ----
    {
        double2 v[NA];
        if (p.src_mode != 0) {
            const double2* src = p.src + (p.src_mode == 2 ? elem_off : 0ull) + tile_base;
#pragma unroll
            for (int i = 0; i < NA; ++i)
                v[i] = (tid + i * nthreads < valid_amps) ? __ldcs(src + TQ_IO_GOFF(i)) : make_double2(0.0, 0.0);
        } else {
#pragma unroll
            for (int i = 0; i < NA; ++i) v[i] = make_double2((tile_base == 0 && tid == 0 && i == 0) ? 1.0 : 0.0, 0.0);
        }
        for (int i = tid; i < 3 * min(n_run, kWinSlots); i += nthreads)
            s_win[i] = __ldg(reinterpret_cast<const uint2*>(p.windows) + i);
        stage_ops(0);
#pragma unroll
        for (int i = 0; i < NA; ++i)
            if (tid + i * nthreads < tile_amps) amp[TQ_IO_SLOT(i)] = v[i];
    }

    // ---- 2. register windows ----
    Amps a;
    uint32_t slot_t = 0, ws0 = 0, ws1 = 0, ws2 = 0, ws3 = 0;
    double acc = 0.0;            // this thread's share of the energy (expectation windows)
    bool regs_dirty = false;     // registers hold amplitudes that shared memory does not have yet
    // entering window w: flush the previous window's registers if they changed, (re)stage ops if needed, compute this
    // thread's layout (slot_t, ws*, ctx) and load its 16 amplitudes.  Defines op_begin / op_end / ctx in scope.
#define TQ_ENTER_WINDOW(w)                                                                                         \
    uint2 h0, h1, h2;                                                                                              \
    if ((w) < kWinSlots) {                                                                                         \
        if ((w) == 0) __syncthreads(); /* headers (and the tile) staged above */                                   \
        h0 = s_win[3 * (w)]; h1 = s_win[3 * (w) + 1]; h2 = s_win[3 * (w) + 2];                                     \
    } else {                                                                                                       \
        const uint2* wraw = reinterpret_cast<const uint2*>(p.windows + (w));                                       \
        h0 = __ldg(wraw); h1 = __ldg(wraw + 1); h2 = __ldg(wraw + 2);                                              \
    }                                                                                                              \
    const uint32_t wpos4 = h0.x;                                                                                   \
    const uint64_t tpos8 = (uint64_t)h0.y | ((uint64_t)h1.x << 32); /* thread bits 0..7 */                         \
    const bool read_only = (h1.y >> 24) & kWinFlagReadOnly;        /* tpos[11] */                                  \
    const int op_begin = (int)h2.x, op_end = (int)h2.y;                                                            \
    if ((w) > 0) {                                                                                                 \
        __syncthreads(); /* every thread has finished the previous window (its loads and its staged ops) */        \
        if (regs_dirty && active) {                                                                                \
            _Pragma("unroll") for (int r = 0; r < NA; ++r) amp[TQ_SLOT(r)] = a[r];                                 \
        }                                                                                                          \
    }                                                                                                              \
    regs_dirty = !read_only;                                                                                       \
    /* ops beyond the staged range (long passes only): restage from this window on (uniform branch) */             \
    if (op_end > staged_end) {                                                                                     \
        if ((w) == 0) __syncthreads();                                                                             \
        stage_ops(op_begin);                                                                                       \
    }                                                                                                              \
    __syncthreads();                                                                                               \
    uint32_t jt = 0;                                                                                               \
    _Pragma("unroll") for (int i = 0; i < 8; ++i)                                                                  \
        if (i < n_tbits && ((tid >> i) & 1)) jt |= 1u << ((uint32_t)(tpos8 >> (8 * i)) & 0xffu);                   \
    const uint64_t ctx = tile_base | (jt < (uint32_t)valid_amps ? TQ_PHYS(jt) : 0u);                               \
    slot_t = swz(jt);                                                                                              \
    ws0 = swz(1u << (wpos4 & 0xffu));                                                                              \
    ws1 = swz(1u << ((wpos4 >> 8) & 0xffu));                                                                       \
    ws2 = swz(1u << ((wpos4 >> 16) & 0xffu));                                                                      \
    ws3 = swz(1u << (wpos4 >> 24));                                                                                \
    if (active) {                                                                                                  \
        _Pragma("unroll") for (int r = 0; r < NA; ++r) a[r] = amp[TQ_SLOT(r)];                                     \
    }

    // ---- 2a. gate windows ----
    for (int w = 0; w < p.n_gate_windows; ++w) {
        TQ_ENTER_WINDOW(w)
        if (active) {
            for (int o = op_begin - staged_begin; o < op_end - staged_begin; ++o) {
                const WinOp wo = s_wops[o];
                const double2* m = s_mat + o * kMatStride;
                const int fc = wo.w0 & 0xff;
                const int qsel = (wo.w0 >> 16) & 0xff;
                switch (fc) {
                case F_U2 + 0: g_u2<0, 1>(a, m); break;
                case F_U2 + 1: g_u2<0, 2>(a, m); break;
                case F_U2 + 2: g_u2<0, 3>(a, m); break;
                case F_U2 + 3: g_u2<1, 2>(a, m); break;
                case F_U2 + 4: g_u2<1, 3>(a, m); break;
                case F_U2 + 5: g_u2<2, 3>(a, m); break;
                case F_U1 + 0: g_u1<0>(a, m); break;
                case F_U1 + 1: g_u1<1>(a, m); break;
                case F_U1 + 2: g_u1<2>(a, m); break;
                case F_U1 + 3: g_u1<3>(a, m); break;
                case F_D1 + 0: g_d1<0>(a, m); break;
                case F_D1 + 1: g_d1<1>(a, m); break;
                case F_D1 + 2: g_d1<2>(a, m); break;
                case F_D1 + 3: g_d1<3>(a, m); break;
                case F_D1_OUT: g_scale(a, ((ctx >> qsel) & 1ull) ? m[3] : m[0]); break;
                case F_CX_OW + 0: g_x_if<0>(a, (bool)((ctx >> qsel) & 1ull)); break;
                case F_CX_OW + 1: g_x_if<1>(a, (bool)((ctx >> qsel) & 1ull)); break;
                case F_CX_OW + 2: g_x_if<2>(a, (bool)((ctx >> qsel) & 1ull)); break;
                case F_CX_OW + 3: g_x_if<3>(a, (bool)((ctx >> qsel) & 1ull)); break;
                default:
                    if (DM) {
                        const int rb = (wo.w0 >> 8) & 0xf, rb2 = (wo.w0 >> 12) & 0xf;
                        if (fc == F_DEPOL1) exec_depol1(a, rb, rb2, wo.fixed);
                        else if (fc == F_DEPOL2) exec_depol2(a, rb, rb2, wo.fixed);
                    }
                    break;
                }
            }
        }
    }
    // ---- 2b. expectation windows (read-only) ----
    if (!DM && p.exp_mode == 1) {
        for (int w = p.n_gate_windows; w < p.n_windows; ++w) {
            TQ_ENTER_WINDOW(w)
            if (active) {
                for (int o = op_begin - staged_begin; o < op_end - staged_begin; ++o) {
                    const WinOp wo = s_wops[o];
                    const double2* m = s_mat + o * kMatStride;
                    const int fc = wo.w0 & 0xff, rb = (wo.w0 >> 8) & 0xf;
                    if (fc == F_EXPC) acc += exec_expc<false>(a, ctx, rb, m);
                    else if (fc == F_EXPC_IMAG) acc += exec_expc<true>(a, ctx, rb, m);
                    else acc += g_expd(a, ctx, m, reinterpret_cast<const double2*>(p.eterms) + wo.t + 2);
                }
            }
        }
    }
    // registers -> shared memory (final layout of the pass)
    if (regs_dirty) {
        __syncthreads();
        if (active) {
#pragma unroll
            for (int r = 0; r < NA; ++r) amp[TQ_SLOT(r)] = a[r];
        }
    }
    __syncthreads();

    // ---- 4. write back ----
    if (p.dst) {
        double2* dst = p.dst + elem_off + tile_base;
#pragma unroll
        for (int i = 0; i < NA; ++i)
            if (tid + i * nthreads < valid_amps) dst[TQ_IO_GOFF(i)] = amp[TQ_IO_SLOT(i)];
    }

    // ---- 3. expectation of the Hamiltonian terms that are local to this pass ----
    if (p.exp_mode != 0) {
        if (p.exp_mode == 1) {  // groups that flip more than kRegBits qubits: partner amplitudes via shared memory
            // terms staged in shared memory (re-using the op staging area: 8 KiB = 256 terms)
            ExpTerm* s_terms = reinterpret_cast<ExpTerm*>(s_mat);
            const int cap = (int)((kOpSlots * kMatStride * sizeof(double2)) / sizeof(ExpTerm));
            for (int g = 0; g < p.n_groups; ++g) {
                const ExpGroup grp = p.groups[g];
                const uint32_t xs = swz(grp.xlocal);
                for (int t0 = grp.term_begin; t0 < grp.term_end; t0 += cap) {
                    const int nt = min(cap, grp.term_end - t0);
                    __syncthreads();
                    for (int i = tid; i < nt; i += nthreads) s_terms[i] = p.terms[t0 + i];
                    __syncthreads();
                    for (int j = tid < nar ? tid : valid_amps; j < valid_amps; j += nar) {
                        const uint32_t sj = swz(j);
                        const double2 v = amp[sj], bq = amp[sj ^ xs];
                        const double px = bq.x * v.x + bq.y * v.y;  // conj(psi[j ^ x]) * psi[j]
                        const double py = bq.x * v.y - bq.y * v.x;
                        double fre = 0.0, fim = 0.0;
                        for (int tt = 0; tt < nt; ++tt) {
                            const ExpTerm term = s_terms[tt];
                            const int par = (__popc((uint32_t)j & term.zlocal) + __popcll(tile_base & term.zphys)) & 1;
                            fre += par ? -term.wre : term.wre;
                            fim += par ? -term.wim : term.wim;
                        }
                        acc += fre * px - fim * py;
                    }
                }
            }
        } else {
            // (with extra staging warps -- PassParams::arith_threads -- the entries are still summed by the plan's own threads
            // in the plan's order: sharing them out over four warps needs a summation order defined over 128 virtual
            // threads for EVERY launch shape, and that cost the batched 8-qubit shape 18 % when it was tried)
            for (int e = tid < nar ? tid : p.n_hent; e < p.n_hent; e += nar) {
                const HEntry h = p.hent[e];
                const double2 ar = amp[swz(h.r)], ac = amp[swz(h.c)];
                const double px = ar.x * ac.x + ar.y * ac.y;  // conj(psi_r) * psi_c
                const double py = ar.x * ac.y - ar.y * ac.x;
                acc += h.re * px - h.im * py;
            }
        }
        const double total = block_sum(acc, s_red, tid, nar);
        if (tid == 0) p.partial[(size_t)b * p.partial_ld + p.partial_off + tile] = total;
    }
}

// =================================================================================================================
// Tensor-core pass kernel (pure states, tiles of >= 2^9 amplitudes): the windows of tq_plan.h "DMMA windows".
// A thread holds NR = 32 doubles: component c = lane & 1 (real / imaginary part) of the 32 amplitudes that differ in
// the window's register qubits; lane bit 1 = QL, lane bits 2..4 and the warp index = untouched tile positions.
// =================================================================================================================
#include "tq_mma_dev.cuh"

#define TQ_SEL5(i, v0, v1, v2, v3, v4) ((((i) & 1) ? (v0) : 0u) ^ (((i) & 2) ? (v1) : 0u) ^ (((i) & 4) ? (v2) : 0u) ^ (((i) & 8) ? (v3) : 0u) ^ (((i) & 16) ? (v4) : 0u))

template <bool TABLE, bool XCHG>   // XCHG: the write-back is the qubit exchange of a sharded state (own instantiation,
                                   // so that the ordinary kernel's register allocation stays as it is)
__global__ void __launch_bounds__(kMaxThreads, TQ_MIN_BLOCKS) tile_pass_mma_kernel(const PassParams p_in) {
    PassParams p_own;
    if (TABLE) p_own = p_in.table[blockIdx.x];
    const PassParams& p = TABLE ? p_own : p_in;
    const uint32_t cta = TABLE ? 0u : blockIdx.x;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tile_amps = 1 << p.k;   // k >= 9: k_eff == k, every thread of every warp is active
    double2* amp = reinterpret_cast<double2*>(smem_raw);
    double* ampd = reinterpret_cast<double*>(smem_raw);
    double2* s_mat = amp + tile_amps;
    WinOp* s_wops = reinterpret_cast<WinOp*>(s_mat + kOpSlots * kMatStride);
    double* s_red = reinterpret_cast<double*>(s_wops + kOpSlots);
    MmaWindowDev* s_win = reinterpret_cast<MmaWindowDev*>(s_red + 32);           // kWinSlots headers

    const int tid = threadIdx.x, nthreads = blockDim.x;
#ifdef TQ_TRACE   // debugging aid: phase timestamps of a few CTAs (scratch builds only)
    __shared__ unsigned long long s_tr[40];
    int n_tr = 0;
    unsigned smid_; asm("mov.u32 %0, %%smid;" : "=r"(smid_));
#define TQ_TR() do { if (tid == 0 && n_tr < 40) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); s_tr[n_tr++] = t_; } } while (0)
#else
#define TQ_TR() do {} while (0)
#endif
    TQ_TR();
    const int lane = tid & 31, warp = tid >> 5;
    const int comp = lane & 1;
    const bool l1 = (lane >> 1) & 1;
    const uint32_t ntiles = 1u << p.n_nl;
    const uint32_t tile = cta & (ntiles - 1u);
    const uint32_t b = cta >> p.n_nl;

    uint64_t tile_base = 0;
    for (int i = 0; i < p.n_nl; ++i) tile_base |= (uint64_t)((tile >> i) & 1u) << p.nonlocal[i];
    // I/O layout as in tile_pass_kernel: amplitude i of thread t is tile index t + i * nthreads (16 per thread); the
    // physical offsets of the thread part and of the four per-bit strides come precomputed from the host
    const uint32_t io_slot = swz(tid);
    const uint32_t iw0 = swz(nthreads), iw1 = swz(nthreads << 1), iw2 = swz(nthreads << 2), iw3 = swz(nthreads << 3);
    const uint32_t io_goff = __ldg(p.io_goff + tid);
    const uint32_t ig0 = p.io_stride[0], ig1 = p.io_stride[1], ig2 = p.io_stride[2], ig3 = p.io_stride[3];
    const uint64_t elem_off = (uint64_t)b << p.nbits;
    const double2* my_mats = p.mats + (size_t)b * p.n_mats * kMatStride;
    if (p.fused_prep)   // single-tile plans: this CTA is the only one of its element and evaluates its block matrices itself
        fused_prep(p, b, tid, nthreads, s_mat, const_cast<double2*>(my_mats));
    const int n_run = (p.exp_mode == 1) ? p.n_windows : p.n_gate_windows;

    // B-fragment coordinates of this lane: B[k = lane & 3][n = lane >> 2]; n = (QL', c', RX'), k = (QL, c)
    const int g = lane >> 2;
    const int brow = (g >> 2) | ((g & 1) << 1);            // staged matrix row: index bit 0 = QL, bit 1 = RX
    const int bcol = (lane >> 1) & 1;                      // column for RX = 0; RX = 1 adds 2
    const bool bsame = ((g >> 1) & 1) == comp;             // output and input component agree -> real part
    const long long bneg = ((g >> 1) & 1) ? 0ll : (long long)(1ull << 63);   // output = real part: -Im

    int staged_begin = 0, staged_end = 0;
    auto stage_ops = [&](int first) {
        staged_begin = first;
        staged_end = min(first + kOpSlots, p.n_wops);
        for (int i = tid; i < (staged_end - staged_begin) * kMatStride; i += nthreads) {
            const int oi = i >> 4, e = i & 15;
            WinOp wo = p.wops[staged_begin + oi];
            const int code = wo.w0 & 0xff, mode = (wo.w0 >> 12) & 0xf;
            if (e == 0) {
                wo.w0 = (wo.w0 & ~0xffu) | (uint32_t)flat_code_mma(wo.w0);
                s_wops[oi] = wo;
            }
            if (code >= M_EXPC) {
                if (e < (code == M_EXPC ? 9 : code == M_EXPT ? 16 : 2))   // header + cA[16] | D[32] | 16 class counts
                    s_mat[oi * kMatStride + e] = reinterpret_cast<const double2*>(p.eterms)[(size_t)wo.t + e];
            } else if (code == M_U2) {
                // expand to a 4x4 with index bit 0 = QL, bit 1 = RX
                const double2* M = my_mats + (size_t)wo.t * kMatStride;
                const int r = e >> 2, c = e & 3;
                double2 v = make_double2(0.0, 0.0);
                if (mode == 0) v = M[e];
                else if (mode == 1) v = M[((((r & 1) << 1) | (r >> 1)) << 2) | ((c & 1) << 1) | (c >> 1)];
                else if (mode == 2) { if ((r & 1) == (c & 1)) v = M[(r >> 1) * 2 + (c >> 1)]; }
                else if (mode == 3) { if ((r >> 1) == (c >> 1)) v = M[(r & 1) * 2 + (c & 1)]; }
                else { if (e == 0 || e == 3) v = M[e]; }
                // kMmaFlagSwapOut: the outputs trade places -> row r goes to the row with its two index bits swapped
                const int dst = ((wo.w0 >> 24) & kMmaFlagSwapOut) ? (((((r & 1) << 1) | (r >> 1)) << 2) | c) : e;
                s_mat[oi * kMatStride + dst] = v;
            }
        }
    };

    // ---- 1. stage the tile, the window headers and the first ops ----
    {
        double2 v[16];
        if (p.direct) {
            // nothing to stage: every window of this pass reads the state itself
        } else if (p.src_mode != 0) {
            const double2* src = p.src + (p.src_mode == 2 ? elem_off : 0ull) + tile_base;
            if (p.in_mask == ~0ull) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __ldcs(src + TQ_IO_GOFF(i));
            } else {   // known zeros of a state grown from |0...0> are not read
                const bool tile_dead = (tile_base & ~p.in_mask) != 0;
                const uint32_t dead = (uint32_t)~p.in_mask;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    v[i] = (tile_dead || (TQ_IO_GOFF(i) & dead)) ? make_double2(0.0, 0.0) : __ldcs(src + TQ_IO_GOFF(i));
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = make_double2((tile_base == 0 && tid == 0 && i == 0) ? 1.0 : 0.0, 0.0);
        }
        for (int i = tid; i < kWinU4 * min(n_run, kWinSlots); i += nthreads)
            reinterpret_cast<uint4*>(s_win)[i] = __ldg(reinterpret_cast<const uint4*>(p.mwindows) + i);
        stage_ops(0);
        if (!p.direct) {
#pragma unroll
            for (int i = 0; i < 16; ++i) amp[TQ_IO_SLOT(i)] = v[i];
        }
    }
    __syncthreads();

    // ---- 2. windows ----
    Regs a;
    double acc = 0.0;
    uint32_t slot_rest = 0;
    uint64_t ctx = 0;
    // entering window w: (re)stage ops if needed, compute this thread's layout and load its 32 doubles
    auto enter = [&](int w) -> const MmaWindowDev* {
        const MmaWindowDev* hdr = (w < kWinSlots) ? s_win + w : nullptr;
        if (!hdr) {   // very long passes: header from global memory into slot 0 (windows run in order)
            __syncthreads();
            if (tid < kWinU4) reinterpret_cast<uint4*>(s_win)[tid] = __ldg(reinterpret_cast<const uint4*>(p.mwindows + w) + tid);
            __syncthreads();
            hdr = s_win;
        }
        if (hdr->op_end > staged_end) {   // CTA-uniform: restage from this window on
            __syncthreads();
            stage_ops(hdr->op_begin);
            __syncthreads();
        }
        // this thread's part of the tile index (lane bits 1..4 and the warp bits): shared-memory slot and physical
        // index bits.  ctx holds QL as it sits on entry: gate windows never read QL through ctx, expectation windows
        // never move it.
        slot_rest = (((lane >> 2) & 1) ? hdr->gslot[0] : 0u) ^ (((lane >> 3) & 1) ? hdr->gslot[1] : 0u) ^
                    (((lane >> 4) & 1) ? hdr->gslot[2] : 0u);
        ctx = tile_base | ((uint64_t)((lane >> 2) & 1) << hdr->gphys[0]) | ((uint64_t)((lane >> 3) & 1) << hdr->gphys[1]) |
              ((uint64_t)((lane >> 4) & 1) << hdr->gphys[2]) | ((uint64_t)l1 << hdr->qlphys);
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (i < p.k - 9 && ((warp >> i) & 1)) {
                slot_rest ^= hdr->wslot[i];
                ctx |= 1ull << hdr->wphys[i];
            }
        if (p.direct) {
            // straight from global memory: amplitude index = ctx | the register bits' physical bits; lanes 2i, 2i+1
            // read the two halves of one amplitude and lane bits 1..3 are qubits 0..2 -> 256 contiguous bytes per load
            const unsigned char* base = reinterpret_cast<const unsigned char*>(p.src + elem_off);
            const uint32_t t = ((uint32_t)ctx << 4) | ((uint32_t)comp << 3);
            const uint32_t x0 = 16u << hdr->rphys[0], x1 = 16u << hdr->rphys[1], x2 = 16u << hdr->rphys[2],
                           x3 = 16u << hdr->rphys[3], x4 = 16u << hdr->rphys[4];
            const uint32_t lo[4] = {0u, x0, x1, x0 ^ x1};
            const uint32_t hi[8] = {t, t ^ x2, t ^ x3, t ^ x2 ^ x3, t ^ x4, t ^ x4 ^ x2, t ^ x4 ^ x3, t ^ x4 ^ x3 ^ x2};
#pragma unroll
            for (int r = 0; r < NR; ++r) a[r] = __ldg(reinterpret_cast<const double*>(base + (hi[r >> 2] ^ lo[r & 3])));
            return hdr;
        }
        if (p.use_dead && (warp & hdr->dead_wbits)) return hdr;   // this warp's share of the tile is all zeros
        // byte offset of register r's double: ((slot_t ^ xor of its bits' slots) << 4) | comp << 3, as three-input XORs
        const uint32_t t = ((slot_rest ^ (l1 ? hdr->qslot : 0u)) << 4) | ((uint32_t)comp << 3);
        const uint32_t x0 = (uint32_t)hdr->rslot[0] << 4, x1 = (uint32_t)hdr->rslot[1] << 4, x2 = (uint32_t)hdr->rslot[2] << 4,
                       x3 = (uint32_t)hdr->rslot[3] << 4, x4 = (uint32_t)hdr->rslot[4] << 4;
        const uint32_t lo[4] = {0u, x0, x1, x0 ^ x1};
        const uint32_t hi[8] = {t, t ^ x2, t ^ x3, t ^ x2 ^ x3, t ^ x4, t ^ x4 ^ x2, t ^ x4 ^ x3, t ^ x4 ^ x3 ^ x2};
#pragma unroll
        for (int r = 0; r < NR; ++r)
            a[r] = *reinterpret_cast<const double*>(smem_raw + (hi[r >> 2] ^ lo[r & 3]));
        return hdr;
    };

    TQ_TR();
    // ---- 2a. gate windows ----
    for (int w = 0; w < p.n_gate_windows; ++w) {
        if (w < kWinSlots && (s_win[w].flags & kWinFlagReadOnly)) continue;   // layout-only window (expectation-only pass)
        const MmaWindowDev* hdr = enter(w);
        if (hdr->flags & kWinFlagReadOnly) continue;
        __syncthreads();   // everyone holds its entry data: the tile may be overwritten from here on
        TQ_TR();
        // a circuit started from |0...0>: warps whose warp-index bits select a qubit nothing has touched yet hold zeros,
        // stay zeros through the window (it does not act on that qubit) and own a region of the tile that is all zeros
        // before and after: they only keep the barriers company
        const bool idle = p.use_dead && (warp & hdr->dead_wbits);
        const int o_end = idle ? 0 : hdr->op_end - staged_begin;
        // op word and B fragment of op o (prefetched one op ahead, so the loads run under the previous block's DMMAs)
        auto fetch = [&](int o, uint32_t& w0, double& b0, double& b1) {
            w0 = s_wops[o].w0;
            const double2* m = s_mat + o * kMatStride;
            double2 u0 = m[brow * 4 + bcol], u1 = m[brow * 4 + bcol + 2];
            if ((w0 & 0xff) == FM_SCAL) {
                const double2 d = ((ctx >> ((w0 >> 16) & 0xff)) & 1ull) ? m[3] : m[0];
                const double2 z = make_double2(0.0, 0.0);
                u0 = (brow == bcol) ? d : z;
                u1 = (brow == (bcol | 2)) ? d : z;
            }
            // real 8x8 form: same component -> Re, re<-im -> -Im, im<-re -> +Im (sign flipped on the integer pipe)
            b0 = bsame ? u0.x : __longlong_as_double(__double_as_longlong(u0.y) ^ bneg);
            b1 = bsame ? u1.x : __longlong_as_double(__double_as_longlong(u1.y) ^ bneg);
        };
        uint32_t w0n = 0;
        double b0n = 0.0, b1n = 0.0;
        int o = hdr->op_begin - staged_begin;
        if (o < o_end) fetch(o, w0n, b0n, b1n);
        for (; o < o_end; ++o) {
            const uint32_t w0 = w0n;
            const double b0 = b0n, b1 = b1n;
            if (o + 1 < o_end) fetch(o + 1, w0n, b0n, b1n);
            const int fc = w0 & 0xff;
            const int qsel = (w0 >> 16) & 0xff;
            if (fc <= FM_SCAL) {   // compare chain (CTA-uniform), cheaper than an indirect branch per block
                const uint32_t dead = p.use_dead ? ((w0 >> 25) & 0x1fu) : 0u;   // flags bits 1..5
                if (fc == FM_U2 + 1) m_u2<1>(a, b0, b1, dead);
                else if (fc == FM_U2 + 2) m_u2<2>(a, b0, b1, dead);
                else if (fc == FM_U2 + 3) m_u2<3>(a, b0, b1, dead);
                else if (fc == FM_U2 + 4) m_u2<4>(a, b0, b1, dead);
                else m_u2<0>(a, b0, b1, dead);   // FM_U2 + 0 and FM_SCAL
            } else {
                switch (fc) {
                case FM_SWAP + 0: m_swapql<0>(a, l1); break;
                case FM_SWAP + 1: m_swapql<1>(a, l1); break;
                case FM_SWAP + 2: m_swapql<2>(a, l1); break;
                case FM_SWAP + 3: m_swapql<3>(a, l1); break;
                case FM_SWAP + 4: m_swapql<4>(a, l1); break;
                case FM_CXO + 0: m_cx_out<0>(a, (bool)((ctx >> qsel) & 1ull)); break;
                case FM_CXO + 1: m_cx_out<1>(a, (bool)((ctx >> qsel) & 1ull)); break;
                case FM_CXO + 2: m_cx_out<2>(a, (bool)((ctx >> qsel) & 1ull)); break;
                case FM_CXO + 3: m_cx_out<3>(a, (bool)((ctx >> qsel) & 1ull)); break;
                case FM_CXO + 4: m_cx_out<4>(a, (bool)((ctx >> qsel) & 1ull)); break;
                default: break;
                }
            }
        }
        TQ_TR();
        if (!idle) {
            const uint32_t t = ((slot_rest ^ (l1 ? hdr->qslot_out : 0u)) << 4) | ((uint32_t)comp << 3);
            const uint32_t x0 = (uint32_t)hdr->rslot_out[0] << 4, x1 = (uint32_t)hdr->rslot_out[1] << 4,
                           x2 = (uint32_t)hdr->rslot_out[2] << 4, x3 = (uint32_t)hdr->rslot_out[3] << 4,
                           x4 = (uint32_t)hdr->rslot_out[4] << 4;
            const uint32_t lo[4] = {0u, x0, x1, x0 ^ x1};
            const uint32_t hi[8] = {t, t ^ x2, t ^ x3, t ^ x2 ^ x3, t ^ x4, t ^ x4 ^ x2, t ^ x4 ^ x3, t ^ x4 ^ x3 ^ x2};
#pragma unroll
            for (int r = 0; r < NR; ++r) *reinterpret_cast<double*>(smem_raw + (hi[r >> 2] ^ lo[r & 3])) = a[r];
        }
        __syncthreads();   // the tile is complete in shared memory again
        TQ_TR();
    }

    // ---- write back (shared memory holds the final tile) ----
    if (XCHG && p.dst && p.xchg_shift) {
        // sharded state: the top bits of the shard index name the rank that owns the amplitude after the exchange -- the
        // tile goes straight into the peers' buffers (16-byte stores over NVLink), no separate all-to-all pass
        const uint32_t low = (1u << p.xchg_shift) - 1u;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t idx = (uint32_t)tile_base | TQ_IO_GOFF(i);
            p.xchg_peer[idx >> p.xchg_shift][p.xchg_self | (idx & low)] = amp[TQ_IO_SLOT(i)];
        }
    } else if (p.dst) {
        double2* dst = p.dst + elem_off + tile_base;
#pragma unroll
        for (int i = 0; i < 16; ++i) dst[TQ_IO_GOFF(i)] = amp[TQ_IO_SLOT(i)];
    }

    // ---- 2b. expectation windows (read-only: no barriers between them) ----
    if (p.exp_mode == 1) {
        for (int w = p.n_gate_windows; w < p.n_windows; ++w) {
            TQ_TR();
            const MmaWindowDev* hdr = enter(w);
#ifdef TQ_TRACE
            if (a[0] == 1.2345e-300) acc += 1.0;   // force the loads to complete before the timestamp
#endif
            TQ_TR();
            const int o_end = hdr->op_end - staged_begin;
            if (hdr->flags & kWinFlagGenericDiag) {
                // diagonal terms with signs from the thread's index: a window of its own, so that this code (and the
                // registers of its Walsh-Hadamard transform) stays out of the loop below
                for (int o = hdr->op_begin - staged_begin; o < o_end; ++o) {
                    const WinOp wo = s_wops[o];
                    const double2* m = s_mat + o * kMatStride;
                    if ((wo.w0 & 0xff) == FM_EXPD) {   // register bit 4 counts as a bit outside the window (qsel)
                        const double2* terms = reinterpret_cast<const double2*>(p.eterms) + wo.t + 2;
                        acc += m_expd_half<0>(a, ctx, m, terms);
                        acc += m_expd_half<1>(a, ctx | (1ull << ((wo.w0 >> 16) & 0xff)), m, terms);
                    } else if ((wo.w0 & 0xff) == FM_EXPT) {
                        const double* D = reinterpret_cast<const double*>(m);
                        double s4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                        for (int r = 0; r < NR; ++r) s4[r & 3] = fma(a[r] * a[r], D[r], s4[r & 3]);
                        acc += (s4[0] + s4[1]) + (s4[2] + s4[3]);
                    }
                }
                continue;
            }
            for (int o = hdr->op_begin - staged_begin; o < o_end; ++o) {
                const WinOp wo = s_wops[o];
                const double2* m = s_mat + o * kMatStride;
                if ((wo.w0 & 0xff) == FM_EXPC) {
                    const double* cA = reinterpret_cast<const double*>(m + 1);
                    const double* cB = ((wo.w0 >> 12) & 1)
                                           ? reinterpret_cast<const double*>(reinterpret_cast<const double2*>(p.eterms) + wo.t + 9)
                                           : nullptr;
                    const double sum = exec_m_expc(a, (int)(wo.w0 >> 24), cA, cB, comp != 0);
                    const uint64_t zphys = (uint64_t)__double_as_longlong(m[0].x);
                    acc += (__popcll(ctx & zphys) & 1) ? -sum : sum;
                } else {   // FM_EXPT: diagonal terms inside the window, signed-weight table
                    const double* D = reinterpret_cast<const double*>(m);
                    double s4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                    for (int r = 0; r < NR; ++r) s4[r & 3] = fma(a[r] * a[r], D[r], s4[r & 3]);
                    acc += (s4[0] + s4[1]) + (s4[2] + s4[3]);
                }
            }
        }
    }

    // ---- 3. Hamiltonian terms evaluated from shared memory (flip masks wider than a window / sparse entries) ----
    if (p.exp_mode != 0) {
        if (p.exp_mode == 1) {
            ExpTerm* s_terms = reinterpret_cast<ExpTerm*>(s_mat);
            const int cap = (int)((kOpSlots * kMatStride * sizeof(double2)) / sizeof(ExpTerm));
            for (int gi = 0; gi < p.n_groups; ++gi) {
                const ExpGroup grp = p.groups[gi];
                const uint32_t xs = swz(grp.xlocal);
                for (int t0 = grp.term_begin; t0 < grp.term_end; t0 += cap) {
                    const int nt = min(cap, grp.term_end - t0);
                    __syncthreads();
                    for (int i = tid; i < nt; i += nthreads) s_terms[i] = p.terms[t0 + i];
                    __syncthreads();
                    for (int j = tid; j < tile_amps; j += nthreads) {
                        const uint32_t sj = swz(j);
                        const double2 v = amp[sj], bq = amp[sj ^ xs];
                        const double px = bq.x * v.x + bq.y * v.y;
                        const double py = bq.x * v.y - bq.y * v.x;
                        double fre = 0.0, fim = 0.0;
                        for (int tt = 0; tt < nt; ++tt) {
                            const ExpTerm term = s_terms[tt];
                            const int par = (__popc((uint32_t)j & term.zlocal) + __popcll(tile_base & term.zphys)) & 1;
                            fre += par ? -term.wre : term.wre;
                            fim += par ? -term.wim : term.wim;
                        }
                        acc += fre * px - fim * py;
                    }
                }
            }
        } else {
            for (int e = tid; e < p.n_hent; e += nthreads) {
                const HEntry h = p.hent[e];
                const double2 ar = amp[swz(h.r)], ac = amp[swz(h.c)];
                const double px = ar.x * ac.x + ar.y * ac.y;
                const double py = ar.x * ac.y - ar.y * ac.x;
                acc += h.re * px - h.im * py;
            }
        }
        const double total = block_sum(acc, s_red, tid, nthreads);
        if (tid == 0) p.partial[(size_t)b * p.partial_ld + p.partial_off + tile] = total;
    }
}

// =================================================================================================================
// Expectation-only passes in ONE persistent launch.  Every "direct" expectation-only pass of a plan (windows whose lane bits
// 1..3 are qubits 0..2: registers loaded straight from the state, no tile staging) becomes a sub-pass of this kernel:
//   * window headers, op words and coefficient tables of all sub-passes are staged in shared memory once per CTA, not
//     once per tile;
//   * a CTA walks the (element, tile) pairs t = blockIdx.x, blockIdx.x + gridDim.x, ... and, for each, runs every
//     sub-pass on ITS tile number t (each sub-pass has its own tiling of the state).  The resident CTAs therefore work
//     on one or two batch elements at a time: the element is read from HBM once and the other sub-passes find it in L2
//     (separate launches would stream the whole batch once per pass);
//   * no barriers except the per-tile sum (double-buffered reduction slots).
// Same arithmetic, same per-(element, pass, tile) partial sums in the same slots as tile_pass_mma_kernel's expectation
// windows: results are bit-identical to the one-launch-per-pass path (TQ_DIRECT_KERNEL=0).
// =================================================================================================================
__global__ void __launch_bounds__(kMaxThreads, TQ_MIN_BLOCKS) expect_direct_kernel(const __grid_constant__ DirectParams dp) {
    __shared__ double2 s_mat[kOpSlots * kMatStride];
    __shared__ WinOp s_wops[kOpSlots];
    __shared__ MmaWindowDev s_win[kWinSlots];
    __shared__ double s_red[32];
    __shared__ int s_wbase[kMaxDirectSub + 1], s_obase[kMaxDirectSub + 1];

    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int comp = lane & 1;
    const bool l1 = (lane >> 1) & 1;

    // ---- stage every sub-pass once ----
    if (tid == 0) {
        int wb = 0, ob = 0;
        for (int s = 0; s < dp.n_sub; ++s) {
            s_wbase[s] = wb;
            s_obase[s] = ob;
            wb += dp.sub[s].n_windows - dp.sub[s].n_gate_windows;
            ob += dp.sub[s].n_wops;
        }
        s_wbase[dp.n_sub] = wb;
        s_obase[dp.n_sub] = ob;
    }
    __syncthreads();
    for (int s = 0; s < dp.n_sub; ++s) {
        const PassParams& p = dp.sub[s];
        const int nw = p.n_windows - p.n_gate_windows;
        for (int i = tid; i < kWinU4 * nw; i += nthreads)
            reinterpret_cast<uint4*>(s_win + s_wbase[s])[i] =
                __ldg(reinterpret_cast<const uint4*>(p.mwindows + p.n_gate_windows) + i);
        for (int i = tid; i < p.n_wops * kMatStride; i += nthreads) {
            const int oi = i >> 4, e = i & 15;
            WinOp wo = p.wops[oi];
            const int code = wo.w0 & 0xff;
            if (e == 0) {
                wo.w0 = (wo.w0 & ~0xffu) | (uint32_t)flat_code_mma(wo.w0);
                s_wops[s_obase[s] + oi] = wo;
            }
            if (code >= M_EXPC && e < (code == M_EXPC ? 9 : code == M_EXPT ? 16 : 2))
                s_mat[(s_obase[s] + oi) * kMatStride + e] = reinterpret_cast<const double2*>(p.eterms)[(size_t)wo.t + e];
        }
    }
    __syncthreads();

    const int n_nl = dp.sub[0].n_nl, nbits = dp.sub[0].nbits, k = dp.sub[0].k;
    const uint32_t ntiles = 1u << n_nl;
    const uint32_t total = (uint32_t)dp.batch << n_nl;
    int flip = 0;
    for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
        const uint32_t tile = t & (ntiles - 1u), b = t >> n_nl;
        const uint64_t elem_off = (uint64_t)b << nbits;
        for (int s = 0; s < dp.n_sub; ++s) {
            const PassParams& p = dp.sub[s];
            uint64_t tile_base = 0;
            for (int i = 0; i < n_nl; ++i) tile_base |= (uint64_t)((tile >> i) & 1u) << p.nonlocal[i];
            const unsigned char* base = reinterpret_cast<const unsigned char*>(p.src + elem_off);
            double acc = 0.0;
            for (int w = s_wbase[s]; w < s_wbase[s + 1]; ++w) {
                const MmaWindowDev* hdr = s_win + w;
                uint64_t ctx = tile_base | ((uint64_t)((lane >> 2) & 1) << hdr->gphys[0]) |
                               ((uint64_t)((lane >> 3) & 1) << hdr->gphys[1]) | ((uint64_t)((lane >> 4) & 1) << hdr->gphys[2]) |
                               ((uint64_t)l1 << hdr->qlphys);
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (i < k - 9 && ((warp >> i) & 1)) ctx |= 1ull << hdr->wphys[i];
                Regs a;
                {
                    const uint32_t tt = ((uint32_t)ctx << 4) | ((uint32_t)comp << 3);
                    const uint32_t x0 = 16u << hdr->rphys[0], x1 = 16u << hdr->rphys[1], x2 = 16u << hdr->rphys[2],
                                   x3 = 16u << hdr->rphys[3], x4 = 16u << hdr->rphys[4];
                    const uint32_t lo[4] = {0u, x0, x1, x0 ^ x1};
                    const uint32_t hi[8] = {tt, tt ^ x2, tt ^ x3, tt ^ x2 ^ x3, tt ^ x4, tt ^ x4 ^ x2, tt ^ x4 ^ x3, tt ^ x4 ^ x3 ^ x2};
#pragma unroll
                    for (int r = 0; r < NR; ++r) a[r] = __ldg(reinterpret_cast<const double*>(base + (hi[r >> 2] ^ lo[r & 3])));
                }
                const int o_begin = s_obase[s] + hdr->op_begin, o_end = s_obase[s] + hdr->op_end;
                if (hdr->flags & kWinFlagGenericDiag) {
                    for (int o = o_begin; o < o_end; ++o) {
                        const WinOp wo = s_wops[o];
                        const double2* m = s_mat + o * kMatStride;
                        if ((wo.w0 & 0xff) == FM_EXPD) {
                            const double2* terms = reinterpret_cast<const double2*>(p.eterms) + wo.t + 2;
                            acc += m_expd_half<0>(a, ctx, m, terms);
                            acc += m_expd_half<1>(a, ctx | (1ull << ((wo.w0 >> 16) & 0xff)), m, terms);
                        } else if ((wo.w0 & 0xff) == FM_EXPT) {
                            const double* D = reinterpret_cast<const double*>(m);
                            double s4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                            for (int r = 0; r < NR; ++r) s4[r & 3] = fma(a[r] * a[r], D[r], s4[r & 3]);
                            acc += (s4[0] + s4[1]) + (s4[2] + s4[3]);
                        }
                    }
                    continue;
                }
                for (int o = o_begin; o < o_end; ++o) {
                    const WinOp wo = s_wops[o];
                    const double2* m = s_mat + o * kMatStride;
                    if ((wo.w0 & 0xff) == FM_EXPC) {
                        const double* cA = reinterpret_cast<const double*>(m + 1);
                        const double* cB = ((wo.w0 >> 12) & 1)
                                               ? reinterpret_cast<const double*>(reinterpret_cast<const double2*>(p.eterms) + wo.t + 9)
                                               : nullptr;
                        const double sum = exec_m_expc(a, (int)(wo.w0 >> 24), cA, cB, comp != 0);
                        const uint64_t zphys = (uint64_t)__double_as_longlong(m[0].x);
                        acc += (__popcll(ctx & zphys) & 1) ? -sum : sum;
                    } else {   // FM_EXPT
                        const double* D = reinterpret_cast<const double*>(m);
                        double s4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                        for (int r = 0; r < NR; ++r) s4[r & 3] = fma(a[r] * a[r], D[r], s4[r & 3]);
                        acc += (s4[0] + s4[1]) + (s4[2] + s4[3]);
                    }
                }
            }
            // deterministic CTA sum (same tree as block_sum); two sets of slots alternate so that a fast warp cannot
            // overwrite what thread 0 is still adding up
            double* red = s_red + (flip ? 16 : 0);
            flip ^= 1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
            if (lane == 0) red[warp] = acc;
            __syncthreads();
            if (tid == 0) {
                double tot = 0.0;
                const int nwarps = (nthreads + 31) >> 5;
                for (int w = 0; w < nwarps; ++w) tot += red[w];
                p.partial[(size_t)b * p.partial_ld + p.partial_off + tile] = tot;
            }
        }
    }
}

// Four threads per (batch element, fused block): thread c carries column c of the block's 4x4 matrix.  A gate mixes ROWS, so
// the columns never meet and each one sees exactly the arithmetic of eval_block_matrix (same results, bit for bit); the next
// gate's descriptor and angle are fetched while the current one is applied (the loop is a chain of dependent loads otherwise).
__global__ void __launch_bounds__(128) prep_matrices_kernel(const MatDesc* __restrict__ descs, const MatGate* __restrict__ prog,
                                                            int n_mats, int batch, const double* __restrict__ params,
                                                            int ld_params, const uint8_t* __restrict__ codes,
                                                            int ld_codes, double2* __restrict__ mats) {
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 2, col = threadIdx.x & 3;
    if (idx >= n_mats * batch) return;
    const int b = idx / n_mats, mi = idx - b * n_mats;
    const MatDesc md = descs[mi];
    double2 M[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) M[r] = make_double2(r == col ? 1.0 : 0.0, 0.0);
    auto angle_of = [&](const MatGate& g) {
        return (g.kind <= MG_RZ && g.pidx >= 0) ? params[(size_t)b * ld_params + g.pidx] : g.fixed;
    };
    MatGate g{};
    double theta = 0.0;
    if (md.begin < md.end) { g = prog[md.begin]; theta = angle_of(g); }
    for (int gi = md.begin; gi < md.end; ++gi) {
        const MatGate cur = g;
        const double th = theta;
        if (gi + 1 < md.end) { g = prog[gi + 1]; theta = angle_of(g); }
        int kind = cur.kind;
        if (kind == MG_CX) {  // control = lq: swap the two rows with the control bit set
            const int ra = cur.lq == 0 ? 1 : 2;
            const double2 t = M[ra]; M[ra] = M[3]; M[3] = t;
            continue;
        }
        if (kind == MG_PAULI_SLOT) {
            const int code = (codes[(size_t)b * ld_codes + cur.pidx] >> (int)cur.fixed) & 3;
            if (code == 0) continue;
            kind = MG_X + code - 1;
        }
        double2 g00 = make_double2(0.0, 0.0), g01 = g00, g10 = g00, g11 = g00;
        if (kind <= MG_RZ) {
            double sn, cs;
            sincos(0.5 * th, &sn, &cs);
            if (kind == MG_RX) { g00.x = cs; g01.y = sn; g10.y = sn; g11.x = cs; }
            else if (kind == MG_RY) { g00.x = cs; g01.x = sn; g10.x = -sn; g11.x = cs; }
            else { g00.x = cs; g00.y = sn; g11.x = cs; g11.y = -sn; }
        } else if (kind == MG_X) { g01.x = 1.0; g10.x = 1.0; }
        else if (kind == MG_Y) { g01.y = -1.0; g10.y = 1.0; }
        else { g00.x = 1.0; g11.x = -1.0; }
        if (cur.lq == 0) {
#pragma unroll
            for (int r = 0; r < 4; r += 2) {
                const double2 x0 = M[r], x1 = M[r + 1];
                M[r] = cfma(g01, x1, cmul(g00, x0));
                M[r + 1] = cfma(g11, x1, cmul(g10, x0));
            }
        } else {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const double2 x0 = M[r], x1 = M[r + 2];
                M[r] = cfma(g01, x1, cmul(g00, x0));
                M[r + 2] = cfma(g11, x1, cmul(g10, x0));
            }
        }
    }
    double2* out = mats + (size_t)idx * kMatStride;
    if (md.nq == 2) {
#pragma unroll
        for (int r = 0; r < 4; ++r) out[r * 4 + col] = M[r];
    } else if (col < 2) {  // one-qubit block: the 2x2 matrix in entries 0..3
        out[col] = M[0];
        out[2 + col] = M[1];
    }
}

// out[b] = sum of row b of `partial` (n entries, row stride ld): one CTA per element, fixed summation order (thread t adds
// entries t, t + 256, ... in order, then the same shuffle tree / warp order as block_sum): re-runs are bit-identical
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ partial, int ld, int n,
                                                               double* __restrict__ out, int batch) {
    __shared__ double s_red[32];
    const int b = blockIdx.x;
    if (b >= batch) return;
    const double* row = partial + (size_t)b * ld;
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += row[i];
    const double total = block_sum(v, s_red, threadIdx.x, blockDim.x);
    if (threadIdx.x == 0) out[b] = total;
}

__global__ void __launch_bounds__(kMaxThreads) dm_expect_kernel(const double2* __restrict__ rho, int n,
                                                                const HEntry* __restrict__ hent, int n_hent,
                                                                double* __restrict__ out) {
    __shared__ double s_red[32];
    const double2* r = rho + ((size_t)blockIdx.x << (2 * n));
    double acc = 0.0;
    for (int e = threadIdx.x; e < n_hent; e += blockDim.x) {
        const HEntry h = hent[e];
        const double2 v = r[(size_t)h.c + ((size_t)h.r << n)];  // rho[c][r]
        acc += h.re * v.x - h.im * v.y;
    }
    const double total = block_sum(acc, s_red, threadIdx.x, blockDim.x);
    if (threadIdx.x == 0) out[blockIdx.x] = total;
}

// ---- FP64 peak microbenchmark (bench.py's roofline denominator for the tensor-core work; MEASURED_PEAKS.json has none) ----
__global__ void __launch_bounds__(256) fp64_peak_kernel(int which, int iters, double* sink) {
    const int lane = threadIdx.x & 31;
    if (which == 0) {   // DMMA: 8 independent accumulator pairs per warp, b-operand fixed
        double d[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) d[i] = 1e-3 * (lane + i);
        const double a = 1.0 + 1e-9 * lane, b = 1.0 - 1e-9 * lane;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) dmma884(d[i], d[i + 1], a, b, d[i], d[i + 1]);
        }
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += d[i];
        if (s == 123.456) sink[0] = s;
    } else {            // DFMA: 16 independent chains per thread
        double d[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) d[i] = 1e-3 * (lane + i);
        const double a = 1.0 - 1e-9, b = 1e-12 * lane;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) d[i] = fma(d[i], a, b);
        }
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += d[i];
        if (s == 123.456) sink[0] = s;
    }
}

}  // namespace

double fp64_peak_run(int which, int n_sms, float* ms_out) {
    const int iters = 4096, blocks = n_sms * 4, threads = 256;
    double* sink = nullptr;
    if (cudaMalloc(&sink, 8) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    fp64_peak_kernel<<<blocks, threads>>>(which, iters, sink);
    cudaEventRecord(e1);
    const cudaError_t err = cudaEventSynchronize(e1);
    cudaEventElapsedTime(ms_out, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) return -1.0;
    const double warps = (double)blocks * threads / 32.0;
    // DMMA m8n8k4: 8 * 8 * 4 FMA = 512 flop per warp instruction; DFMA: 2 flop per lane
    return which == 0 ? warps * iters * 8.0 * 512.0 : warps * 32.0 * iters * 16.0 * 2.0;
}

size_t tile_pass_smem_bytes(int k_eff, int k, int lead) {
    const size_t n_hi = (size_t)1 << (k - lead);
    return ((size_t)16 << k_eff) + kOpSlots * (kMatStride * sizeof(double2) + sizeof(WinOp)) + 32 * sizeof(double) +
           kWinSlots * sizeof(MmaWindowDev) + n_hi * sizeof(uint32_t);   // MmaWindowDev (64 B) >= Window (24 B)
}

cudaError_t tile_pass_configure() {
    const void* kernels[] = {(const void*)tile_pass_kernel<false, false>, (const void*)tile_pass_kernel<true, false>,
                             (const void*)tile_pass_kernel<false, true>, (const void*)tile_pass_mma_kernel<false, false>,
                             (const void*)tile_pass_mma_kernel<true, false>, (const void*)tile_pass_mma_kernel<false, true>};
    for (const void* k : kernels) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

void launch_tile_pass(const PassParams& p, int batch, int threads, bool density, cudaStream_t stream) {
    const unsigned grid = (unsigned)batch << p.n_nl;
    const size_t smem = tile_pass_smem_bytes(p.k_eff, p.k, p.lead);
    if (p.mwindows && p.xchg_shift) tile_pass_mma_kernel<false, true><<<grid, threads, smem, stream>>>(p);
    else if (p.mwindows) tile_pass_mma_kernel<false, false><<<grid, threads, smem, stream>>>(p);
    else if (density) tile_pass_kernel<true, false><<<grid, threads, smem, stream>>>(p);
    else tile_pass_kernel<false, false><<<grid, threads, smem, stream>>>(p);
}

void launch_tile_pass_table(const PassParams* table_dev, int n, int threads, size_t smem, bool mma, cudaStream_t stream) {
    PassParams head{};
    head.table = table_dev;
    if (mma) tile_pass_mma_kernel<true, false><<<n, threads, smem, stream>>>(head);
    else tile_pass_kernel<false, true><<<n, threads, smem, stream>>>(head);
}

void launch_expect_direct(const DirectParams& dp, int n_ctas, int threads, cudaStream_t stream) {
    expect_direct_kernel<<<n_ctas, threads, 0, stream>>>(dp);
}

void launch_prep_matrices(const MatDesc* descs, const MatGate* prog, int n_mats, int batch, const double* params,
                          int ld_params, const uint8_t* codes, int ld_codes, double2* mats, cudaStream_t stream) {
    const int total = n_mats * batch;
    if (total <= 0) return;
    prep_matrices_kernel<<<(total * 4 + 127) / 128, 128, 0, stream>>>(descs, prog, n_mats, batch, params, ld_params, codes,
                                                                   ld_codes, mats);
}

void launch_reduce_partials(const double* partial, int ld, int n, double* out, int batch, cudaStream_t stream) {
    reduce_partials_kernel<<<batch, n >= 1024 ? 256 : 64, 0, stream>>>(partial, ld, n, out, batch);
}

void launch_dm_expect(const double2* rho, int n, const HEntry* hent, int n_hent, double* out, int batch,
                      cudaStream_t stream) {
    dm_expect_kernel<<<batch, kMaxThreads, 0, stream>>>(rho, n, hent, n_hent, out);
}

}  // namespace tq
