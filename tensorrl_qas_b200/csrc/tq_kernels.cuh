// tq_kernels.cuh -- sm_100a kernels of libtqsim (batched complex128 statevector / density-matrix passes).
//
// One kernel does the heavy lifting: tile_pass_kernel.  A CTA owns one tile (2^k amplitudes that differ only in
// the pass's k local qubits) of one batch element:
//   1. stage the tile in shared memory with coalesced 16-byte loads (runs of 2^lead consecutive amplitudes),
//      bank-swizzled so that any later redistribution is conflict free;
//   2. run the pass's register windows: every thread holds 16 amplitudes (4 window qubits) in registers and applies
//      all fused blocks of the window there (dense 4x4 / 2x2 / diagonal blocks whose per-element matrices come from
//      prep_matrices_kernel, CNOTs, exact depolarising channels); shared memory is touched only to switch windows;
//   3. optionally evaluate the Hamiltonian terms whose flip masks are local (one deterministic partial per tile);
//   4. write the tile back.
// HBM traffic per pass is one read + one write of the state no matter how many gates the pass fused.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tq_plan.h"

namespace tq {

constexpr int kMaxTileBits = 12;
constexpr int kMaxShardRanks = 8;   // ranks one state can be sharded over (one NVSwitch box)
constexpr int kMaxThreads = 256; // 2^(kMaxTileBits - kRegBits)

struct ExpGroup {   // one X/Y flip mask of the Hamiltonian, local to the pass
    uint32_t xlocal;
    int32_t term_begin, term_end;
    int32_t pad;
};
struct ExpTerm {    // coefficient already multiplied by i^{#Y}
    uint64_t zphys;   // Z/Y bits on physical (non-tile) positions -> sign from the tile's base index
    uint32_t zlocal;  // Z/Y bits on tile positions
    uint32_t pad;
    double wre, wim;
};
struct HEntry {     // non-zero of the symmetrised Hamiltonian, upper triangle, off-diagonals pre-doubled
    uint32_t r, c;
    double re, im;
};

struct PassParams {
    const PassParams* table;   // table launches (launch_tile_pass_table): per-CTA parameters, everything else unused
    int nbits, k, k_eff, lead, n_nl;   // k = real tile bits, k_eff = max(k, kRegBits)
    uint8_t local[16];
    uint8_t nonlocal[32];
    const double2* src;
    int src_mode;  // 0: |0...0>, 1: one shared initial vector, 2: per-element state buffer
    double2* dst;  // nullptr: tile is not written back
    const Window* windows;
    const MmaWindowDev* mwindows;   // non-null: tensor-core pass (tile_pass_mma_kernel), `windows` unused
    const uint32_t* io_goff;        // tensor-core pass: physical offset of tile index t, t < threads (global table)
    uint32_t io_stride[4];          // ... and of tile indices threads << i
    // single-tile plans may fold prep_matrices_kernel into the pass (one launch per energy call): the CTA of element b
    // evaluates its own block matrices into mats[b] first
    int fused_prep;
    const MatDesc* descs;
    const MatGate* prog;
    int n_prog;                     // gates in `prog` (0 = unknown: the program is read from global memory gate by gate)
    int arith_threads;              // > 0: the CTA was launched with MORE threads than the plan's (latency path: the extra
                                    // warps only help with staging); the energy is accumulated by the first arith_threads
                                    // threads in the plan's own order, so the result does not depend on the launch shape
    const double* params;
    const uint8_t* codes;
    int ld_params, ld_codes;
    uint64_t in_mask;               // tensor-core pass: qubits that can be 1 in the input state (~0: any).  A circuit
                                    // started from |0...0> only populates the qubits its gates have touched so far:
                                    // amplitudes with another bit set are known zeros and are not read (nor were
                                    // they written by the pass before)
    int use_dead;                   // honour MmaWindowDev::dead_wbits (the run started from |0...0>)
    int direct;                     // expectation-only pass: no tile staging, windows load from `src` directly
    int n_windows, n_gate_windows;   // expectation windows follow the gate windows
    const EUnit* eterms;
    const WinOp* wops;
    int n_wops;
    const double2* mats;   // [batch][n_mats][kMatStride] block matrices of this call's elements (prep kernel output)
    int n_mats;
    int exp_mode;  // 0 none, 1 Pauli sum (expectation windows + shared-memory groups), 2 sparse entries (single tile)
    const ExpGroup* groups;
    int n_groups;
    const ExpTerm* terms;
    int n_terms;
    const HEntry* hent;
    int n_hent;
    double* partial;  // partial[b * partial_ld + partial_off + tile]
    int partial_ld, partial_off;
    // single-state sharding (tensor-core passes, one element): the write-back IS the qubit exchange.  Amplitude idx of this
    // rank's shard goes to rank idx >> xchg_shift, at (xchg_self | (idx & low bits)) of that rank's buffer xchg_peer[...]
    // (peer memory over NVLink, or this rank's own second buffer).  xchg_shift == 0: ordinary write-back to `dst`.
    int xchg_shift;
    uint32_t xchg_self;             // this rank's number << xchg_shift
    double2* xchg_peer[kMaxShardRanks];
};

// the direct expectation-only passes of a plan as sub-passes of one persistent launch (expect_direct_kernel)
constexpr int kMaxDirectSub = 4;
constexpr int kDirectOpSlots = 48, kDirectWinSlots = 32;   // = kOpSlots / kWinSlots of tq_kernels.cu: totals over the sub-passes
struct DirectParams {
    int n_sub, batch;
    PassParams sub[kMaxDirectSub];
};
void launch_expect_direct(const DirectParams& dp, int n_ctas, int threads, cudaStream_t stream);

// ---- streaming pass kernel (tq_stream.cu): persistent CTAs, tiles moved by TMA --------------------------------------
constexpr int kStreamMaxSub = 4;      // sub-passes per launch (expectation-only passes share one launch)
constexpr int kStreamThreads = 512;   // two groups of 256 threads, one tile each
constexpr int kStreamWarpSlots = kStreamThreads / 32;   // partial-sum slots per tile of an expectation-only sub-pass (one per warp)
struct StreamTma {                    // one direction of one sub-pass (tq_plan.h StreamLayout, resolved)
    int n_ops;                        // TMA operations per tile; 0 = direction unused
    uint32_t box_bytes;               // bytes per operation = distance of the operations' shared-memory destinations
    uint32_t tile_bytes;              // n_ops * box_bytes: the mbarrier transaction count of a tile load
    uint32_t op_goff[kStreamMaxOps];  // physical amplitude offset of operation i inside the tile
};
struct StreamSub {
    PassParams pp;             // geometry, ops, eterms, mats, partial: as for the pass kernels (its window pointers are unused)
    const StreamWindowDev* swindows;   // the pass's windows resolved for the streaming layouts (stream_window_dev)
    StreamTma in, out;
    uint64_t in_elem_stride;   // amplitudes between consecutive elements of the source (0: one shared initial vector)
    int has_gates;             // the sub-pass has gate windows (block matrices are staged per element)
};
struct alignas(64) StreamParams {
    CUtensorMap map_in[kStreamMaxSub];   // rank-5 views over doubles; dim 0 = 16 doubles whose coordinate is the tile's base
    CUtensorMap map_out;                 // gate passes (n_sub == 1) that write the tile back
    StreamSub sub[kStreamMaxSub];
    int n_sub, batch;
    int stagger_ns;   // gate passes: the second group starts this much later, so that the groups' DMMA phases interleave
    int one_group;    // diagnostic (TQ_STREAM_ONE_GROUP): gate passes run on one group of 256 threads per CTA
    int chain_windows;   // expectation windows of nearest-neighbour chains run as one fused routine (TQ_STREAM_CHAIN, default 1)
    int contiguous;   // 1: CTA c takes a contiguous range of tiles (gate passes: block matrices are staged once per element);
                      // 0: tiles c, c + grid, ... element-major (expectation sub-passes find the element in L2)
};
size_t tile_stream_smem_bytes();
cudaError_t tile_stream_configure();   // opt in to the dynamic shared memory
bool tile_stream_base_ok(cudaError_t* err);   // the shared-memory layout the streaming kernel assumes holds on this device
void launch_tile_stream(const StreamParams& sp, int n_ctas, cudaStream_t stream);

// FP64 peak of the device, measured (tq_fp64_peak): which = 0 mma.sync.m8n8k4.f64 chains as the pass kernels issue them,
// 1 DFMA chains on the FP64 pipe.  Returns the flop executed and the kernel time.
double fp64_peak_run(int which, int n_sms, float* ms_out);

size_t tile_pass_smem_bytes(int k_eff, int k, int lead);
cudaError_t tile_pass_configure();  // opt in to > 48 KiB dynamic shared memory
void launch_tile_pass(const PassParams& p, int batch, int threads, bool density, cudaStream_t stream);
// n CTAs, CTA i runs the single-tile pure-state problem table_dev[i] (one element each): B different circuits per launch
void launch_tile_pass_table(const PassParams* table_dev, int n, int threads, size_t smem, bool mma, cudaStream_t stream);

// block matrices of every (element, fused block): mats[(b * n_mats + m) * kMatStride ...]
void launch_prep_matrices(const MatDesc* descs, const MatGate* prog, int n_mats, int batch, const double* params,
                          int ld_params, const uint8_t* codes, int ld_codes, double2* mats, cudaStream_t stream);

// out[b] = sum_{s < n} partial[b * ld + s], fixed order
void launch_reduce_partials(const double* partial, int ld, int n, double* out, int batch, cudaStream_t stream);

// out[b] = sum_e Re(h_e * rho_b[c_e + (r_e << n)])
void launch_dm_expect(const double2* rho, int n, const HEntry* hent, int n_hent, double* out, int batch,
                      cudaStream_t stream);

}  // namespace tq
