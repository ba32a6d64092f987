// tq_kernels.cu -- see tq_kernels.cuh.  Compiled only for sm_100a.
#include "tq_kernels.cuh"

namespace tq {
namespace {

constexpr int NA = 1 << kRegBits;  // amplitudes per thread

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

// ---- register-window gates.  RB = register bit of the target; all loops are fully unrolled (static indices) ------
template <int RB>
__device__ __forceinline__ void g_rx(Amps& a, double c, double s) {  // exp(+i t/2 X) = [[c, i s], [i s, c]]
#pragma unroll
    for (int m = 0; m < NA / 2; ++m) {
        const int i0 = ((m >> RB) << (RB + 1)) | (m & ((1 << RB) - 1)), i1 = i0 | (1 << RB);
        const double2 a0 = a[i0], a1 = a[i1];
        a[i0] = make_double2(c * a0.x - s * a1.y, c * a0.y + s * a1.x);
        a[i1] = make_double2(c * a1.x - s * a0.y, c * a1.y + s * a0.x);
    }
}
template <int RB>
__device__ __forceinline__ void g_ry(Amps& a, double c, double s) {  // exp(+i t/2 Y) = [[c, s], [-s, c]]
#pragma unroll
    for (int m = 0; m < NA / 2; ++m) {
        const int i0 = ((m >> RB) << (RB + 1)) | (m & ((1 << RB) - 1)), i1 = i0 | (1 << RB);
        const double2 a0 = a[i0], a1 = a[i1];
        a[i0] = make_double2(c * a0.x + s * a1.x, c * a0.y + s * a1.y);
        a[i1] = make_double2(c * a1.x - s * a0.x, c * a1.y - s * a0.y);
    }
}
template <int RB>
__device__ __forceinline__ void g_rz(Amps& a, double c, double s) {  // diag(c + i s, c - i s)
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        const double sj = ((i >> RB) & 1) ? -s : s;
        const double2 v = a[i];
        a[i] = make_double2(c * v.x - sj * v.y, c * v.y + sj * v.x);
    }
}
__device__ __forceinline__ void g_phase(Amps& a, double c, double s) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        const double2 v = a[i];
        a[i] = make_double2(c * v.x - s * v.y, c * v.y + s * v.x);
    }
}
template <int RB>
__device__ __forceinline__ void g_x_if(Amps& a, bool pred) {  // X on RB where pred (per-thread control bit)
#pragma unroll
    for (int m = 0; m < NA / 2; ++m) {
        const int i0 = ((m >> RB) << (RB + 1)) | (m & ((1 << RB) - 1)), i1 = i0 | (1 << RB);
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
template <int RB>
__device__ __forceinline__ void g_y(Amps& a, double sign) {  // Y = [[0, -i], [i, 0]]; sign -1: conj(Y) = -Y
#pragma unroll
    for (int m = 0; m < NA / 2; ++m) {
        const int i0 = ((m >> RB) << (RB + 1)) | (m & ((1 << RB) - 1)), i1 = i0 | (1 << RB);
        const double2 a0 = a[i0], a1 = a[i1];
        a[i0] = make_double2(sign * a1.y, -sign * a1.x);
        a[i1] = make_double2(-sign * a0.y, sign * a0.x);
    }
}
template <int RB>
__device__ __forceinline__ void g_z(Amps& a) {
#pragma unroll
    for (int i = 0; i < NA; ++i)
        if ((i >> RB) & 1) a[i] = make_double2(-a[i].x, -a[i].y);
}
__device__ __forceinline__ void g_neg_if(Amps& a, bool pred) {
    const double f = pred ? -1.0 : 1.0;
#pragma unroll
    for (int i = 0; i < NA; ++i) a[i] = make_double2(f * a[i].x, f * a[i].y);
}
template <int RB>
__device__ __forceinline__ void g_pauli(Amps& a, int code) {
    if (code == 1) g_x_if<RB>(a, true);
    else if (code == 2) g_y<RB>(a, 1.0);
    else if (code == 3) g_z<RB>(a);
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

#define TQ_SLOT(r) (slot_t ^ (((r) & 1) ? ws0 : 0u) ^ (((r) & 2) ? ws1 : 0u) ^ (((r) & 4) ? ws2 : 0u) ^ (((r) & 8) ? ws3 : 0u))

__global__ void __launch_bounds__(kMaxThreads, 2) tile_pass_kernel(const PassParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tile_amps = 1 << p.k_eff;
    const int valid_amps = 1 << p.k;
    double2* amp = reinterpret_cast<double2*>(smem_raw);
    double2* s_trig = amp + tile_amps;                                  // kOpsChunk entries
    WinOp* s_wops = reinterpret_cast<WinOp*>(s_trig + kOpsChunk);       // kOpsChunk entries
    double* s_red = reinterpret_cast<double*>(s_wops + kOpsChunk);      // 32 entries
    uint32_t* hi_off = reinterpret_cast<uint32_t*>(s_red + 32);         // 2^(k - lead) entries

    const int tid = threadIdx.x, nthreads = blockDim.x;
    const uint32_t ntiles = 1u << p.n_nl;
    const uint32_t tile = blockIdx.x & (ntiles - 1u);
    const uint32_t b = blockIdx.x >> p.n_nl;

    uint64_t tile_base = 0;
    for (int i = 0; i < p.n_nl; ++i) tile_base |= (uint64_t)((tile >> i) & 1u) << p.nonlocal[i];

    const int n_hi = 1 << (p.k - p.lead);
    for (int h = tid; h < n_hi; h += nthreads) {
        uint32_t off = 0;
        for (int i = p.lead; i < p.k; ++i) off |= ((uint32_t)(h >> (i - p.lead)) & 1u) << p.local[i];
        hi_off[h] = off;
    }
    __syncthreads();

    const uint64_t elem_off = (uint64_t)b << p.nbits;
    const uint32_t lead_mask = (1u << p.lead) - 1u;

    // ---- 1. stage the tile (coalesced global reads, swizzled shared-memory slots) ----
    if (p.src_mode == 0) {
        for (int j = tid; j < tile_amps; j += nthreads) amp[j] = make_double2(0.0, 0.0);
        __syncthreads();
        if (tile_base == 0 && tid == 0) amp[0] = make_double2(1.0, 0.0);  // swz(0) == 0
    } else {
        const double2* src = p.src + (p.src_mode == 2 ? elem_off : 0ull) + tile_base;
        for (int j = tid; j < tile_amps; j += nthreads)
            amp[swz(j)] = j < valid_amps ? src[hi_off[j >> p.lead] | (j & lead_mask)] : make_double2(0.0, 0.0);
    }

    const double* my_params = p.params ? p.params + (size_t)b * p.ld_params : nullptr;
    const uint8_t* my_codes = p.codes ? p.codes + (size_t)b * p.ld_codes : nullptr;
    const bool active = tid < (tile_amps >> kRegBits);
    const int n_tbits = p.k_eff - kRegBits;

    // ---- 2. register windows ----
    Amps a;
    uint32_t slot_t = 0, ws0 = 0, ws1 = 0, ws2 = 0, ws3 = 0;
    int staged_begin = 0, staged_end = 0;
    for (int w = 0; w < p.n_windows; ++w) {
        const Window* win = p.windows + w;
        const int op_begin = win->op_begin, op_end = win->op_end;
        if (w > 0) {
            __syncthreads();  // every thread has finished the previous window (its loads and its staged ops)
            if (active) {
#pragma unroll
                for (int r = 0; r < NA; ++r) amp[TQ_SLOT(r)] = a[r];
            }
        }
        // stage this window's ops with their cos/sin while no amplitudes are live in registers; the planner keeps
        // a window's op range within kOpsChunk, so a staged chunk always covers whole windows (uniform branch)
        if (op_end > staged_end) {
            staged_begin = op_begin;
            staged_end = min(op_begin + kOpsChunk, p.n_wops);
            for (int i = tid; i < staged_end - staged_begin; i += nthreads) {
                const WinOp wo = p.wops[staged_begin + i];
                s_wops[i] = wo;
                const int code = wo.w0 & 0xff;
                if (code <= W_PHASE) {
                    const double theta = wo.t >= 0 ? my_params[wo.t] : wo.fixed;
                    double s, c;
                    sincos(0.5 * theta, &s, &c);
                    if (((wo.w0 >> 24) & FLAG_CONJ) && code != W_ROT_Y) s = -s;
                    s_trig[i] = make_double2(c, s);
                }
            }
        }
        __syncthreads();
        uint32_t jt = 0;
        uint64_t ctx = tile_base;  // physical index of this thread's amplitudes with the window bits cleared
        for (int i = 0; i < n_tbits; ++i) {
            if ((tid >> i) & 1) {
                const int pos = win->tpos[i];
                jt |= 1u << pos;
                ctx |= 1ull << p.local[pos];
            }
        }
        slot_t = swz(jt);
        ws0 = swz(1u << win->wpos[0]);
        ws1 = swz(1u << win->wpos[1]);
        ws2 = swz(1u << win->wpos[2]);
        ws3 = swz(1u << win->wpos[3]);
        if (active) {
#pragma unroll
            for (int r = 0; r < NA; ++r) a[r] = amp[TQ_SLOT(r)];
            for (int o = op_begin; o < op_end; ++o) {
                const WinOp wo = s_wops[o - staged_begin];
                const double2 cs = s_trig[o - staged_begin];
                const int code = wo.w0 & 0xff, rb = (wo.w0 >> 8) & 0xf, rb2 = (wo.w0 >> 12) & 0xf;
                const int qsel = (wo.w0 >> 16) & 0xff;
                switch (code) {
                case W_ROT_X: TQ_RB4(g_rx, rb, a, cs.x, cs.y); break;
                case W_ROT_Y: TQ_RB4(g_ry, rb, a, cs.x, cs.y); break;
                case W_ROT_Z: TQ_RB4(g_rz, rb, a, cs.x, cs.y); break;
                case W_PHASE: g_phase(a, cs.x, ((ctx >> qsel) & 1ull) ? -cs.y : cs.y); break;
                case W_CX_WW: exec_cx_ww(a, rb, rb2); break;
                case W_CX_OW: TQ_RB4(g_x_if, rb, a, (bool)((ctx >> qsel) & 1ull)); break;
                case W_X: TQ_RB4(g_x_if, rb, a, true); break;
                case W_Y: TQ_RB4(g_y, rb, a, ((wo.w0 >> 24) & FLAG_CONJ) ? -1.0 : 1.0); break;
                case W_Z: TQ_RB4(g_z, rb, a); break;
                case W_Z_OUT: g_neg_if(a, (bool)((ctx >> qsel) & 1ull)); break;
                case W_PAULI: { const int pc = (my_codes[wo.t] >> rb2) & 3; TQ_RB4(g_pauli, rb, a, pc); break; }
                case W_DEPOL1: exec_depol1(a, rb, rb2, wo.fixed); break;
                case W_DEPOL2: exec_depol2(a, rb, rb2, wo.fixed); break;
                default: break;
                }
            }
        }
    }
    // registers -> shared memory (final layout of the pass)
    if (p.n_windows > 0) {
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
        for (int j = tid; j < valid_amps; j += nthreads) dst[hi_off[j >> p.lead] | (j & lead_mask)] = amp[swz(j)];
    }

    // ---- 3. expectation of the Hamiltonian terms that are local to this pass ----
    if (p.exp_mode != 0) {
        double acc = 0.0;
        if (p.exp_mode == 1) {
            // terms staged in shared memory (re-using the op staging area: 8 KiB = 256 terms)
            ExpTerm* s_terms = reinterpret_cast<ExpTerm*>(s_trig);
            const int cap = (int)((kOpsChunk * (sizeof(double2) + sizeof(WinOp))) / sizeof(ExpTerm));
            for (int g = 0; g < p.n_groups; ++g) {
                const ExpGroup grp = p.groups[g];
                const uint32_t xs = swz(grp.xlocal);
                for (int t0 = grp.term_begin; t0 < grp.term_end; t0 += cap) {
                    const int nt = min(cap, grp.term_end - t0);
                    __syncthreads();
                    for (int i = tid; i < nt; i += nthreads) s_terms[i] = p.terms[t0 + i];
                    __syncthreads();
                    for (int j = tid; j < valid_amps; j += nthreads) {
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
            for (int e = tid; e < p.n_hent; e += nthreads) {
                const HEntry h = p.hent[e];
                const double2 ar = amp[swz(h.r)], ac = amp[swz(h.c)];
                const double px = ar.x * ac.x + ar.y * ac.y;  // conj(psi_r) * psi_c
                const double py = ar.x * ac.y - ar.y * ac.x;
                acc += h.re * px - h.im * py;
            }
        }
        const double total = block_sum(acc, s_red, tid, nthreads);
        if (tid == 0) p.partial[(size_t)b * p.partial_ld + p.partial_off + tile] = total;
    }
}

__global__ void reduce_partials_kernel(const double* __restrict__ partial, int ld, int n, double* __restrict__ out,
                                       int batch) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= batch) return;
    const double* row = partial + (size_t)warp * ld;
    double v = 0.0;
    for (int i = lane; i < n; i += 32) v += row[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) out[warp] = v;
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

}  // namespace

size_t tile_pass_smem_bytes(int k_eff, int k, int lead) {
    const size_t n_hi = (size_t)1 << (k - lead);
    return ((size_t)16 << k_eff) + kOpsChunk * (sizeof(double2) + sizeof(WinOp)) + 32 * sizeof(double) +
           n_hi * sizeof(uint32_t);
}

cudaError_t tile_pass_configure() {
    return cudaFuncSetAttribute(tile_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
}

void launch_tile_pass(const PassParams& p, int batch, int threads, cudaStream_t stream) {
    const unsigned grid = (unsigned)batch << p.n_nl;
    tile_pass_kernel<<<grid, threads, tile_pass_smem_bytes(p.k_eff, p.k, p.lead), stream>>>(p);
}

void launch_reduce_partials(const double* partial, int ld, int n, double* out, int batch, cudaStream_t stream) {
    const int threads = 128, warps_per_block = threads / 32;
    reduce_partials_kernel<<<(batch + warps_per_block - 1) / warps_per_block, threads, 0, stream>>>(partial, ld, n, out,
                                                                                                  batch);
}

void launch_dm_expect(const double2* rho, int n, const HEntry* hent, int n_hent, double* out, int batch,
                      cudaStream_t stream) {
    dm_expect_kernel<<<batch, kMaxThreads, 0, stream>>>(rho, n, hent, n_hent, out);
}

}  // namespace tq
