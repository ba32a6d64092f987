// tq_kernels.cu -- see tq_kernels.cuh.  Compiled only for sm_100a.
#include "tq_kernels.cuh"

namespace tq {
namespace {

__device__ __forceinline__ uint32_t insert_zero(uint32_t m, int pos) {
    return ((m >> pos) << (pos + 1)) | (m & ((1u << pos) - 1u));
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

struct TileCtx {
    double2* amp;
    int tile_amps;
    int tid, nthreads;
    uint64_t tile_base;  // physical index bits of the non-local qubits of this tile
};

// exp(+i theta/2 X) = [[c, i s], [i s, c]]
__device__ __forceinline__ void op_rx(const TileCtx& t, int pos, double c, double s) {
    const uint32_t bit = 1u << pos;
    for (uint32_t m = t.tid; m < (uint32_t)(t.tile_amps >> 1); m += t.nthreads) {
        const uint32_t i0 = insert_zero(m, pos), i1 = i0 | bit;
        const double2 a0 = t.amp[i0], a1 = t.amp[i1];
        t.amp[i0] = make_double2(c * a0.x - s * a1.y, c * a0.y + s * a1.x);
        t.amp[i1] = make_double2(c * a1.x - s * a0.y, c * a1.y + s * a0.x);
    }
}

// exp(+i theta/2 Y) = [[c, s], [-s, c]]
__device__ __forceinline__ void op_ry(const TileCtx& t, int pos, double c, double s) {
    const uint32_t bit = 1u << pos;
    for (uint32_t m = t.tid; m < (uint32_t)(t.tile_amps >> 1); m += t.nthreads) {
        const uint32_t i0 = insert_zero(m, pos), i1 = i0 | bit;
        const double2 a0 = t.amp[i0], a1 = t.amp[i1];
        t.amp[i0] = make_double2(c * a0.x + s * a1.x, c * a0.y + s * a1.y);
        t.amp[i1] = make_double2(c * a1.x - s * a0.x, c * a1.y - s * a0.y);
    }
}

// exp(+i theta/2 Z) = diag(c + i s, c - i s); `mask`/`value` select which amplitudes take the conjugate phase
__device__ __forceinline__ void op_rz(const TileCtx& t, int pos, double c, double s) {
    for (uint32_t j = t.tid; j < (uint32_t)t.tile_amps; j += t.nthreads) {
        const double sj = ((j >> pos) & 1u) ? -s : s;
        const double2 a = t.amp[j];
        t.amp[j] = make_double2(c * a.x - sj * a.y, c * a.y + sj * a.x);
    }
}

__device__ __forceinline__ void op_phase_all(const TileCtx& t, double c, double s) {
    for (uint32_t j = t.tid; j < (uint32_t)t.tile_amps; j += t.nthreads) {
        const double2 a = t.amp[j];
        t.amp[j] = make_double2(c * a.x - s * a.y, c * a.y + s * a.x);
    }
}

__device__ __forceinline__ void op_x(const TileCtx& t, int pos) {
    const uint32_t bit = 1u << pos;
    for (uint32_t m = t.tid; m < (uint32_t)(t.tile_amps >> 1); m += t.nthreads) {
        const uint32_t i0 = insert_zero(m, pos), i1 = i0 | bit;
        const double2 a0 = t.amp[i0];
        t.amp[i0] = t.amp[i1];
        t.amp[i1] = a0;
    }
}

// Y = [[0, -i], [i, 0]]; sign = -1 gives conj(Y) = -Y
__device__ __forceinline__ void op_y(const TileCtx& t, int pos, double sign) {
    const uint32_t bit = 1u << pos;
    for (uint32_t m = t.tid; m < (uint32_t)(t.tile_amps >> 1); m += t.nthreads) {
        const uint32_t i0 = insert_zero(m, pos), i1 = i0 | bit;
        const double2 a0 = t.amp[i0], a1 = t.amp[i1];
        t.amp[i0] = make_double2(sign * a1.y, -sign * a1.x);
        t.amp[i1] = make_double2(-sign * a0.y, sign * a0.x);
    }
}

__device__ __forceinline__ void op_z(const TileCtx& t, int pos) {
    for (uint32_t j = t.tid; j < (uint32_t)t.tile_amps; j += t.nthreads) {
        if ((j >> pos) & 1u) {
            const double2 a = t.amp[j];
            t.amp[j] = make_double2(-a.x, -a.y);
        }
    }
}

__device__ __forceinline__ void op_negate_all(const TileCtx& t) {
    for (uint32_t j = t.tid; j < (uint32_t)t.tile_amps; j += t.nthreads) {
        const double2 a = t.amp[j];
        t.amp[j] = make_double2(-a.x, -a.y);
    }
}

__device__ __forceinline__ void op_cnot(const TileCtx& t, int cpos, int tpos) {
    const uint32_t bit = 1u << tpos;
    for (uint32_t m = t.tid; m < (uint32_t)(t.tile_amps >> 1); m += t.nthreads) {
        const uint32_t i0 = insert_zero(m, tpos);
        if ((i0 >> cpos) & 1u) {
            const uint32_t i1 = i0 | bit;
            const double2 a0 = t.amp[i0];
            t.amp[i0] = t.amp[i1];
            t.amp[i1] = a0;
        }
    }
}

__device__ __forceinline__ void op_pauli_code(const TileCtx& t, int pos, int code) {
    if (code == 1) op_x(t, pos);
    else if (code == 2) op_y(t, pos, 1.0);
    else if (code == 3) op_z(t, pos);
}

// exact 1-qubit depolarising channel on the (row bit, column bit) pair of a density matrix:
// rho -> (1 - 4p/3) rho + (2p/3) Tr_q(rho) (x) I
__device__ __forceinline__ void op_depol1_dm(const TileCtx& t, int pa, int pb, double p) {
    const int lo = pa < pb ? pa : pb, hi = pa < pb ? pb : pa;
    const uint32_t ba = 1u << pa, bb = 1u << pb;
    const double keep = 1.0 - 2.0 * p / 3.0, mixw = 2.0 * p / 3.0, off = 1.0 - 4.0 * p / 3.0;
    for (uint32_t m = t.tid; m < (uint32_t)(t.tile_amps >> 2); m += t.nthreads) {
        const uint32_t i00 = insert_zero(insert_zero(m, lo), hi);
        const double2 r00 = t.amp[i00], r11 = t.amp[i00 | ba | bb];
        const double2 r10 = t.amp[i00 | ba], r01 = t.amp[i00 | bb];
        t.amp[i00] = make_double2(keep * r00.x + mixw * r11.x, keep * r00.y + mixw * r11.y);
        t.amp[i00 | ba | bb] = make_double2(keep * r11.x + mixw * r00.x, keep * r11.y + mixw * r00.y);
        t.amp[i00 | ba] = make_double2(off * r10.x, off * r10.y);
        t.amp[i00 | bb] = make_double2(off * r01.x, off * r01.y);
    }
}

// exact 2-qubit depolarising channel: rho -> (1 - 16p/15) rho + (4p/15) Tr_ab(rho) (x) I_4
__device__ __forceinline__ void op_depol2_dm(const TileCtx& t, int ra, int rb, int ca, int cb, double p) {
    int s[4] = {ra, rb, ca, cb};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3 - i; ++j)
            if (s[j] > s[j + 1]) { const int tmp = s[j]; s[j] = s[j + 1]; s[j + 1] = tmp; }
    const double alpha = 1.0 - 16.0 * p / 15.0, beta = 4.0 * p / 15.0;
    for (uint32_t m = t.tid; m < (uint32_t)(t.tile_amps >> 4); m += t.nthreads) {
        uint32_t base = m;
#pragma unroll
        for (int i = 0; i < 4; ++i) base = insert_zero(base, s[i]);
        double2 tr = make_double2(0.0, 0.0);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const uint32_t idx = base | ((x & 1) ? (1u << ra) | (1u << ca) : 0u) | ((x & 2) ? (1u << rb) | (1u << cb) : 0u);
            const double2 v = t.amp[idx];
            tr.x += v.x;
            tr.y += v.y;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t idx = base | ((r & 1) ? (1u << ra) : 0u) | ((r & 2) ? (1u << rb) : 0u) |
                                     ((c & 1) ? (1u << ca) : 0u) | ((c & 2) ? (1u << cb) : 0u);
                double2 v = t.amp[idx];
                v.x *= alpha;
                v.y *= alpha;
                if (r == c) { v.x += beta * tr.x; v.y += beta * tr.y; }
                t.amp[idx] = v;
            }
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

__global__ void __launch_bounds__(kMaxThreads) tile_pass_kernel(const PassParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tile_amps = 1 << p.k;
    double2* amp = reinterpret_cast<double2*>(smem_raw);
    double2* s_trig = amp + tile_amps;
    DevOp* s_ops = reinterpret_cast<DevOp*>(s_trig + kOpsChunk);
    double* s_red = reinterpret_cast<double*>(s_ops + kOpsChunk);
    uint32_t* hi_off = reinterpret_cast<uint32_t*>(s_red + 32);

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

    // ---- stage the tile ----
    if (p.src_mode == 0) {
        for (int j = tid; j < tile_amps; j += nthreads) amp[j] = make_double2(0.0, 0.0);
        if (tile_base == 0 && tid == 0) amp[0] = make_double2(1.0, 0.0);
    } else {
        const double2* src = p.src + (p.src_mode == 2 ? elem_off : 0ull) + tile_base;
        for (int j = tid; j < tile_amps; j += nthreads) amp[j] = src[hi_off[j >> p.lead] | (j & lead_mask)];
    }

    TileCtx t{amp, tile_amps, tid, nthreads, tile_base};
    const double* my_params = p.params ? p.params + (size_t)b * p.ld_params : nullptr;
    const uint8_t* my_codes = p.codes ? p.codes + (size_t)b * p.ld_codes : nullptr;

    // ---- gates, staged in chunks together with their cos/sin ----
    for (int base = 0; base < p.n_ops; base += kOpsChunk) {
        const int cnt = min(kOpsChunk, p.n_ops - base);
        __syncthreads();
        for (int i = tid; i < cnt; i += nthreads) {
            const DevOp o = p.ops[base + i];
            s_ops[i] = o;
            if (o.op <= OP_RZ_NL) {
                const double theta = o.t >= 0 ? my_params[o.t] : o.fixed;
                double s, c;
                sincos(0.5 * theta, &s, &c);
                if ((o.flags & FLAG_CONJ) && o.op != OP_RY) s = -s;
                s_trig[i] = make_double2(c, s);
            }
        }
        __syncthreads();
        for (int i = 0; i < cnt; ++i) {
            const DevOp o = s_ops[i];
            const double2 cs = s_trig[i];
            switch (o.op) {
            case OP_RX: op_rx(t, o.a, cs.x, cs.y); break;
            case OP_RY: op_ry(t, o.a, cs.x, cs.y); break;
            case OP_RZ: op_rz(t, o.a, cs.x, cs.y); break;
            case OP_RZ_NL: op_phase_all(t, cs.x, ((tile_base >> o.a) & 1ull) ? -cs.y : cs.y); break;
            case OP_CNOT: op_cnot(t, o.a, o.b); break;
            case OP_CNOT_NL: if ((tile_base >> o.a) & 1ull) op_x(t, o.b); break;
            case OP_X: op_x(t, o.a); break;
            case OP_Y: op_y(t, o.a, (o.flags & FLAG_CONJ) ? -1.0 : 1.0); break;
            case OP_Z: op_z(t, o.a); break;
            case OP_Z_NL: if ((tile_base >> o.a) & 1ull) op_negate_all(t); break;
            case OP_PAULI1: op_pauli_code(t, o.a, my_codes[o.t] & 3); break;
            case OP_PAULI2: {
                const int code = my_codes[o.t];
                op_pauli_code(t, o.a, code & 3);
                __syncthreads();
                op_pauli_code(t, o.b, (code >> 2) & 3);
                break;
            }
            case OP_DEPOL1_DM: op_depol1_dm(t, o.a, o.b, o.fixed); break;
            case OP_DEPOL2_DM: op_depol2_dm(t, o.a & 0xff, (o.a >> 8) & 0xff, o.b & 0xff, (o.b >> 8) & 0xff, o.fixed); break;
            default: break;
            }
            __syncthreads();
        }
    }
    __syncthreads();

    // ---- write back ----
    if (p.dst) {
        double2* dst = p.dst + elem_off + tile_base;
        for (int j = tid; j < tile_amps; j += nthreads) dst[hi_off[j >> p.lead] | (j & lead_mask)] = amp[j];
    }

    // ---- expectation of the Hamiltonian terms that are local to this pass ----
    if (p.exp_mode != 0) {
        double acc = 0.0;
        if (p.exp_mode == 1) {
            for (int g = 0; g < p.n_groups; ++g) {
                const ExpGroup grp = p.groups[g];
                for (int j = tid; j < tile_amps; j += nthreads) {
                    const double2 a = amp[j], bq = amp[j ^ grp.xlocal];
                    const double px = bq.x * a.x + bq.y * a.y;  // conj(psi[j ^ x]) * psi[j]
                    const double py = bq.x * a.y - bq.y * a.x;
                    double fre = 0.0, fim = 0.0;
                    for (int tt = grp.term_begin; tt < grp.term_end; ++tt) {
                        const ExpTerm term = p.terms[tt];
                        const int par = (__popc((uint32_t)j & term.zlocal) + __popcll(tile_base & term.zphys)) & 1;
                        fre += par ? -term.wre : term.wre;
                        fim += par ? -term.wim : term.wim;
                    }
                    acc += fre * px - fim * py;
                }
            }
        } else {
            for (int e = tid; e < p.n_hent; e += nthreads) {
                const HEntry h = p.hent[e];
                const double2 ar = amp[h.r], ac = amp[h.c];
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

size_t tile_pass_smem_bytes(int k, int lead) {
    const size_t n_hi = (size_t)1 << (k - lead);
    return ((size_t)16 << k) + kOpsChunk * sizeof(double2) + kOpsChunk * sizeof(DevOp) + 32 * sizeof(double) +
           n_hi * sizeof(uint32_t);
}

cudaError_t tile_pass_configure() {
    return cudaFuncSetAttribute(tile_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

void launch_tile_pass(const PassParams& p, int batch, int threads, cudaStream_t stream) {
    const unsigned grid = (unsigned)batch << p.n_nl;
    tile_pass_kernel<<<grid, threads, tile_pass_smem_bytes(p.k, p.lead), stream>>>(p);
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
