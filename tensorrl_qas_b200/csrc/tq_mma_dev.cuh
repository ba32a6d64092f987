// tq_mma_dev.cuh -- device helpers of the tensor-core (DMMA) windows, shared by tile_pass_mma_kernel / expect_direct_kernel
// (tq_kernels.cu) and the streaming pass kernel (tq_stream.cu).  Include INSIDE `namespace tq { namespace {`.
// A thread holds NR = 32 doubles: component c = lane & 1 (real / imaginary part) of the 32 amplitudes that differ in the
// window's register qubits; lane bit 1 = QL, lane bits 2..4 and the warp index = untouched tile positions (tq_plan.h).
#ifndef TQ_MMA_CHAINS
#define TQ_MMA_CHAINS 2   // independent DMMA chains per warp inside a dense block (tq_stream.cu: 4)
#endif
constexpr int NR = 1 << kMmaRegBits;
typedef double Regs[NR];
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
        : "=d"(d0), "=d"(d1)
        : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// 4x4 complex block on (QL, RX) as an 8x8 real product: per register pair (RX = 0, RX = 1) two chained DMMAs;
// A = the registers themselves (row = lane group, k = (QL, re/im) of the lane), B = (b0, b1) from the block's matrix,
// D = the new pair, which the instruction delivers in two ADJACENT registers.  With RX on register bit 0 that is the
// pair itself.  With RX on bit X != 0 the four registers of (bit X, bit 0) are processed together and written back
// with the two bits' roles exchanged -- still in place, no register moves; the planner tracks the relabelling.
// `dead`: register bits whose qubit no gate has populated yet (runs from |0...0>, CTA-uniform): register pairs with such
// a bit set hold zeros and are skipped; if RX itself is dead its RX = 1 inputs are zeros and one DMMA per pair suffices.
template <int X>
__device__ __forceinline__ void m_u2(Regs& a, double b0, double b1, uint32_t dead) {
    if (dead == 0) {   // the common case: a straight run of 64 DMMAs
#if TQ_MMA_CHAINS == 4
        // four independent chains in flight (the DMMAs are volatile asm: they issue in program order, so the second
        // stage of a register quad must not follow its first stage directly)
        if (X == 0) {
#pragma unroll
            for (int p = 0; p < NR; p += 8) {
                double t0, t1, u0, u1, v0, v1, w0, w1;
                dmma884(t0, t1, a[p], b0, 0.0, 0.0);
                dmma884(u0, u1, a[p + 2], b0, 0.0, 0.0);
                dmma884(v0, v1, a[p + 4], b0, 0.0, 0.0);
                dmma884(w0, w1, a[p + 6], b0, 0.0, 0.0);
                dmma884(a[p], a[p + 1], a[p + 1], b1, t0, t1);
                dmma884(a[p + 2], a[p + 3], a[p + 3], b1, u0, u1);
                dmma884(a[p + 4], a[p + 5], a[p + 5], b1, v0, v1);
                dmma884(a[p + 6], a[p + 7], a[p + 7], b1, w0, w1);
            }
        } else {
#pragma unroll
            for (int o = 0; o < NR / 4; o += 2) {
                const int lo = o & ((1 << (X - 1)) - 1), hi = o >> (X - 1);
                const int p00 = (hi << (X + 1)) | (lo << 1), p01 = p00 | 1, p10 = p00 | (1 << X), p11 = p10 | 1;
                const int lo2 = (o + 1) & ((1 << (X - 1)) - 1), hi2 = (o + 1) >> (X - 1);
                const int q00 = (hi2 << (X + 1)) | (lo2 << 1), q01 = q00 | 1, q10 = q00 | (1 << X), q11 = q10 | 1;
                double t0, t1, u0, u1, v0, v1, w0, w1;
                dmma884(t0, t1, a[p00], b0, 0.0, 0.0);
                dmma884(u0, u1, a[p01], b0, 0.0, 0.0);
                dmma884(v0, v1, a[q00], b0, 0.0, 0.0);
                dmma884(w0, w1, a[q01], b0, 0.0, 0.0);
                dmma884(a[p00], a[p01], a[p10], b1, t0, t1);
                dmma884(a[p10], a[p11], a[p11], b1, u0, u1);
                dmma884(a[q00], a[q01], a[q10], b1, v0, v1);
                dmma884(a[q10], a[q11], a[q11], b1, w0, w1);
            }
        }
#else
        if (X == 0) {
#pragma unroll
            for (int p = 0; p < NR; p += 4) {   // two independent chains in flight
                double t0, t1, u0, u1;
                dmma884(t0, t1, a[p], b0, 0.0, 0.0);
                dmma884(u0, u1, a[p + 2], b0, 0.0, 0.0);
                dmma884(a[p], a[p + 1], a[p + 1], b1, t0, t1);
                dmma884(a[p + 2], a[p + 3], a[p + 3], b1, u0, u1);
            }
        } else {
#pragma unroll
            for (int o = 0; o < NR / 4; ++o) {
                // o enumerates the register bits other than 0 and X
                const int lo = o & ((1 << (X - 1)) - 1), hi = o >> (X - 1);
                const int p00 = (hi << (X + 1)) | (lo << 1), p01 = p00 | 1, p10 = p00 | (1 << X), p11 = p10 | 1;
                double t0, t1, u0, u1;
                dmma884(t0, t1, a[p00], b0, 0.0, 0.0);   // bit-0 qubit = 0: inputs RX = 0 / 1 are p00 / p10
                dmma884(u0, u1, a[p01], b0, 0.0, 0.0);   // bit-0 qubit = 1: inputs p01 / p11
                dmma884(a[p00], a[p01], a[p10], b1, t0, t1);   // -> (bit X = 0; bit 0 = RX')
                dmma884(a[p10], a[p11], a[p11], b1, u0, u1);   // -> (bit X = 1; bit 0 = RX')
            }
        }
#endif
        return;
    }
    // early in a run from |0...0>: skip the register pairs that are known zeros
    if (X == 0) {
        const bool x_dead = dead & 1u;
#pragma unroll
        for (int p = 0; p < NR; p += 2) {
            if (p & dead) continue;
            if (x_dead) {
                dmma884(a[p], a[p + 1], a[p], b0, 0.0, 0.0);
            } else {
                double t0, t1;
                dmma884(t0, t1, a[p], b0, 0.0, 0.0);
                dmma884(a[p], a[p + 1], a[p + 1], b1, t0, t1);
            }
        }
    } else {
        const bool x_dead = (dead >> X) & 1u;
        const uint32_t others = dead & ~(1u | (1u << X));
#pragma unroll
        for (int o = 0; o < NR / 4; ++o) {
            const int lo = o & ((1 << (X - 1)) - 1), hi = o >> (X - 1);
            const int p00 = (hi << (X + 1)) | (lo << 1), p01 = p00 | 1, p10 = p00 | (1 << X), p11 = p10 | 1;
            if (p00 & others) continue;
            if (x_dead) {   // inputs with RX = 1 (p10, p11) are zeros: one DMMA per chain, the second chain first (it reads p01)
                dmma884(a[p10], a[p11], a[p01], b0, 0.0, 0.0);
                dmma884(a[p00], a[p01], a[p00], b0, 0.0, 0.0);
            } else {
                double t0, t1, u0, u1;
                dmma884(t0, t1, a[p00], b0, 0.0, 0.0);
                dmma884(u0, u1, a[p01], b0, 0.0, 0.0);
                dmma884(a[p00], a[p01], a[p10], b1, t0, t1);
                dmma884(a[p10], a[p11], a[p11], b1, u0, u1);
            }
        }
    }
}
// QL <-> RX: the registers with RX != (lane's QL bit) cross to lane ^ 2
template <int X>
__device__ __forceinline__ void m_swapql(Regs& a, bool l1) {
#pragma unroll
    for (int p = 0; p < NR / 2; ++p) {
        const int r0 = ((p >> X) << (X + 1)) | (p & ((1 << X) - 1)), r1 = r0 | (1 << X);
        const double send = l1 ? a[r0] : a[r1];
        const double recv = __shfl_xor_sync(kFull, send, 2);
        if (l1) a[r0] = recv;
        else a[r1] = recv;
    }
}
template <int X>
__device__ __forceinline__ void m_cx_out(Regs& a, bool pred) {
#pragma unroll
    for (int p = 0; p < NR / 2; ++p) {
        const int r0 = ((p >> X) << (X + 1)) | (p & ((1 << X) - 1)), r1 = r0 | (1 << X);
        const double a0 = a[r0], a1 = a[r1];
        a[r0] = pred ? a1 : a0;
        a[r1] = pred ? a0 : a1;
    }
}
// expectation class on registers: every lane contributes the products of ITS component; the two lanes of a pair are
// summed by the CTA-wide reduction.  Im(conj(w) v) needs the other component of v: one shuffle per register pair.
// a product the compiler must leave where it is: the state registers do not change between the expectation ops of a
// window, so every a[r] * a[r'] is loop-invariant there, and hoisting them all out of the op loop (31 flip patterns x 16
// pairs) costs far more registers than a thread has
__device__ __forceinline__ double mul_here(double x, double y) {
    double p;
    asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(p) : "d"(x), "d"(y));
    return p;
}
template <int XR>
__device__ __forceinline__ double m_expc(const Regs& a, const double* __restrict__ cA, const double* __restrict__ cB,
                                         bool im_lane) {
    double s[4] = {0.0, 0.0, 0.0, 0.0};   // four independent accumulation chains
    {
        int q = 0;
        double2 c = make_double2(0.0, 0.0);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            if ((r ^ XR) > r) {
                if ((q & 1) == 0) c = *reinterpret_cast<const double2*>(cA + q);   // coefficients two at a time
                s[q & 3] = fma((q & 1) ? c.y : c.x, mul_here(a[r ^ XR], a[r]), s[q & 3]);
                ++q;
            }
        }
    }
    if (cB) {  // CTA-uniform, rare (terms with an odd number of Y factors): Im(conj(w) v) needs the other component of v
        int q = 0;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            if ((r ^ XR) > r) {
                const double vp = __shfl_xor_sync(kFull, a[r], 1);
                const double im = im_lane ? -mul_here(a[r ^ XR], vp) : mul_here(a[r ^ XR], vp);   // w.x v.y  |  -w.y v.x
                s[q & 3] = fma(-__ldg(cB + q), im, s[q & 3]);
                ++q;
            }
        }
    }
    return (s[0] + s[1]) + (s[2] + s[3]);
}
__device__ __forceinline__ double exec_m_expc(const Regs& a, int xr, const double* cA, const double* cB, bool im_lane) {
#define TQ_XC(V) case V: return m_expc<V>(a, cA, cB, im_lane);
    switch (xr) {
        TQ_XC(1) TQ_XC(2) TQ_XC(3) TQ_XC(4) TQ_XC(5) TQ_XC(6) TQ_XC(7) TQ_XC(8) TQ_XC(9) TQ_XC(10) TQ_XC(11)
        TQ_XC(12) TQ_XC(13) TQ_XC(14) TQ_XC(15) TQ_XC(16) TQ_XC(17) TQ_XC(18) TQ_XC(19) TQ_XC(20) TQ_XC(21)
        TQ_XC(22) TQ_XC(23) TQ_XC(24) TQ_XC(25) TQ_XC(26) TQ_XC(27) TQ_XC(28) TQ_XC(29) TQ_XC(30)
    default: return m_expc<31>(a, cA, cB, im_lane);
    }
#undef TQ_XC
}
// diagonal terms over register bits 0..3 of one half (register bit 4 fixed): WHT of 16 squared components, then one
// signed weight sum per class
template <int HALF>
__device__ __forceinline__ double m_expd_half(const Regs& a, uint64_t ctx, const double2* __restrict__ head,
                                              const double2* __restrict__ terms) {
    double n[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) n[r] = a[HALF * 16 + r] * a[HALF * 16 + r];
#pragma unroll
    for (int bitp = 0; bitp < 4; ++bitp)
#pragma unroll
        for (int r = 0; r < 16; ++r)
            if (!((r >> bitp) & 1)) {
                const double x = n[r], y = n[r | (1 << bitp)];
                n[r] = x + y;
                n[r | (1 << bitp)] = x - y;
            }
    const unsigned short* cnt = reinterpret_cast<const unsigned short*>(head);
    double total = 0.0;
    int idx = 0;
#pragma unroll
    for (int zr = 0; zr < 16; ++zr) {
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

enum : int { FM_U2 = 0, FM_SCAL = 5, FM_SWAP = 6, FM_CXO = 11, FM_EXPC = 17, FM_EXPD = 18, FM_EXPT = 19 };

__device__ __forceinline__ int flat_code_mma(uint32_t w0) {
    const int code = w0 & 0xff, rb = (w0 >> 8) & 0xf, rb2 = (w0 >> 12) & 0xf;
    switch (code) {
    case M_U2: return rb2 == 4 ? FM_SCAL : FM_U2 + rb;
    case M_SWAPQL: return FM_SWAP + rb;
    case M_CX_OUT: return FM_CXO + rb;
    case M_EXPC: return FM_EXPC;
    case M_EXPT: return FM_EXPT;
    default: return FM_EXPD;
    }
}
