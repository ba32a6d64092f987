// tq_plan.h -- host-side circuit compiler of libtqsim: gate list -> fused blocks -> tile passes -> register windows.
//
// Reference behaviour being replaced: qulacs applies one gate per full-state pass
// (circuit.update_quantum_state, environments/VQAs/VQE_qulacs.py:83).  Here:
//
//  1. FUSION.  Runs of gates that act on the same one or two qubits are merged into one dense 2x2 / 4x4 block
//     (the transpiled SU(4) bricks of the MPS init circuit collapse to one block each).  A block's matrix depends
//     on the batch element's angles (and, for trajectory noise, on its sampled Pauli codes), so it is described
//     by a small "matrix program" that a prep kernel evaluates per (element, block).  RZ/Z-only runs stay
//     diagonal; a lone CNOT stays a CNOT (its control only needs to be read, not mixed).
//  2. PASSES.  A pass streams the batch of state vectors through shared memory once: every CTA loads one tile of
//     2^k amplitudes (those that differ only in the pass's k "local" qubits), applies the pass's blocks and
//     writes the tile back.  Blocks whose mixing qubits are all local can run in the pass; diagonal action on a
//     non-local qubit (a diagonal block, the control of a CNOT) only needs the tile's fixed bit.  Blocks that act
//     on disjoint qubits commute, so a block that does not fit is deferred with everything that later touches it.
//  3. WINDOWS.  Inside a tile every thread keeps 2^kRegBits amplitudes in registers (the window's qubits); blocks
//     on window qubits run on registers, shared memory is touched only to switch windows.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace tq {

// ---- matrix programs (evaluated per batch element by prep_matrices_kernel) -------------------------------------
enum : int32_t {  // MatGate.kind
    MG_RX = 0, MG_RY = 1, MG_RZ = 2,   // on block qubit lq; angle = params[pidx] or fixed
    MG_CX = 3,                         // control = block qubit lq, target = the other one
    MG_X = 4, MG_Y = 5, MG_Z = 6,      // fixed Paulis on lq
    MG_PAULI_SLOT = 7,                 // Pauli (codes[pidx] >> shift) & 3 on lq (trajectory noise); shift in `fixed`
};
struct MatGate {   // 24 bytes
    int32_t kind, lq, pidx, pad;
    double fixed;
};
struct MatDesc {   // one fused block: gates [begin, end) of the program, 1 or 2 qubits
    int32_t begin, end, nq, diag;   // diag: all gates diagonal -> only entries (0,0), (1,1) are used
};
constexpr int kMatStride = 16;      // complex entries per block matrix in device memory (4x4, row-major)

// ---- tile-level ops (positions are tile-local bit positions unless noted) --------------------------------------
enum : int32_t {
    OP_U2 = 0,        // a = pos(q0), b = pos(q1) (q0 < q1 physical; matrix index bit 0 = q0), t = matrix
    OP_U1 = 1,        // a = pos, t = matrix
    OP_D1 = 2,        // diagonal 1-qubit block: a = pos, t = matrix
    OP_D1_NL = 3,     // a = physical bit (non-local): whole tile times d0 or d1
    OP_CNOT = 4,      // a = control pos, b = target pos
    OP_CNOT_NL = 5,   // a = physical control bit (non-local), b = target pos
    OP_DEPOL1_DM = 6, // a = pos(q), b = pos(q + n), fixed = p
    OP_DEPOL2_DM = 7, // a = pos(qa) | pos(qb) << 8, b = pos(qa + n) | pos(qb + n) << 8, fixed = p
};
enum : int32_t { FLAG_CONJ = 1, FLAG_SWAP = 2 };  // conj: column side of a density matrix; swap: matrix qubit order

struct DevOp {
    int32_t op, a, b, t, flags, pad;
    double fixed;
};

struct Gate {        // user-level gate (tq_set_circuit)
    int32_t kind, q0, q1, pidx;
    double fixed;
};

// ---- register windows -------------------------------------------------------------------------------------------
constexpr int kRegBits = 4;
constexpr int kMinTileBits = kRegBits;   // smaller states are padded with phantom positions
constexpr int kMaxWindowOps = 32;        // ops per window (the kernel stages one window's ops + matrices at a time)

enum : int32_t {    // WinOp code; rb / rb2 are register-bit indices inside the window
    W_U2 = 0,       // rb < rb2: register bits of matrix index bits 0 / 1 (FLAG_SWAP if the block's q0 sits on rb2)
    W_U1 = 1,       // rb
    W_D1 = 2,       // rb
    W_D1_OUT = 3,   // diagonal block on a bit outside the window: qsel = physical bit
    W_CX_WW = 4,    // rb = control, rb2 = target
    W_CX_OW = 5,    // qsel = physical control bit, rb = target
    W_DEPOL1 = 6,   // rb = row bit, rb2 = column bit, fixed = p
    W_DEPOL2 = 7,   // window = {row a, row b, col a, col b}: rb = ra | rb << 2, rb2 = ca | cb << 2
    W_EXPC = 8,     // expectation, flip mask rb != 0 over the register bits: the terms of the group that share the same
                    // Z/Y bits outside the window ("class").  Data (9 units at Pass::eterms[t]): unit 0 = those outside
                    // bits (physical mask), units 1..8 = (cA[q], cB[q]) for the 8 register pairs r < r ^ rb:
                    //   E += sign(ctx) * sum_q (cA[q] Re p_q - cB[q] Im p_q),  p_q = conj(psi[r ^ rb]) psi[r];
                    // rb2 bit 0: some cB != 0
    W_EXPD = 9,     // expectation of all diagonal terms (flip mask 0): units 0..1 at eterms[t] = 16 uint16 counts of the
                    // terms per class zr (= Z bits inside the window), then one unit (outside Z mask, weight) per term,
                    // sorted by class:  E += sum_zr WHT(|psi|^2)[zr] * sum_{t in zr} w_t sign_t(ctx)
};
constexpr int kWinFlagReadOnly = 1;  // Window::tpos[11]: the window does not change the amplitudes
constexpr int kWinFlagGenericDiag = 4;  // MmaWindow (read-only): holds an M_EXPD op (and no M_EXPC): the kernel keeps that code
                                        // path out of the ordinary expectation loop
constexpr int kWinFlagDirect = 2;    // MmaWindow (read-only): lane bits 1..3 are the three lowest qubits, so the thread
                                     // can load its registers straight from global memory, fully coalesced

struct EUnit { uint64_t w[2]; };   // 16 bytes of expectation data (bit patterns of masks / doubles)
struct ExpTermIn { uint64_t z; double wre, wim; };
struct ExpGroupIn { uint64_t x; std::vector<ExpTermIn> terms; };
struct WinOp {      // 16 bytes
    uint32_t w0;    // code | rb << 8 | rb2 << 12 | qsel << 16 | flags << 24
    int32_t t;      // matrix index (W_U2 / W_U1 / W_D1 / W_D1_OUT)
    double fixed;
};
inline uint32_t winop_pack(int code, int rb, int rb2, int qsel, int flags) {
    return (uint32_t)code | ((uint32_t)rb << 8) | ((uint32_t)rb2 << 12) | ((uint32_t)qsel << 16) | ((uint32_t)flags << 24);
}
struct Window {     // 24 bytes, read by the kernel from global memory
    uint8_t wpos[kRegBits];  // tile position of register bit r
    uint8_t tpos[12];        // tile position of thread bit i (first k - kRegBits entries used); the first three
                             // are chosen with independent bank-swizzle vectors (conflict-free exchanges)
    int32_t op_begin, op_end;  // range in Pass::wops
};

// bank swizzle of the shared-memory tile: amplitude j lives in slot j ^ fold(j >> 3), a GF(2)-linear map that
// spreads any three independent tile positions over the eight 16-byte bank groups
constexpr uint8_t kSwizzleVec[13] = {1, 2, 4, 3, 5, 6, 7, 1, 2, 4, 3, 5, 6};
inline uint32_t swizzle_slot(uint32_t j) {
    uint32_t s = j;
    for (int p = 3; p < 13; ++p)
        if ((j >> p) & 1) s ^= kSwizzleVec[p];
    return s;
}

// ---- tensor-core (DMMA) windows -----------------------------------------------------------------------------------
// Pure-state passes with tiles of >= 2^9 amplitudes run their blocks on the FP64 tensor cores (mma.sync m8n8k4.f64).
// A warp owns 2^9 amplitudes; a thread holds 32 doubles: ONE component (lane bit 0: real / imaginary part) of the 32
// amplitudes that differ in the window's five "register qubits" R0..R4.  Lane bit 1 is one more window qubit, "QL";
// lane bits 2..4 and the warp index are six tile positions the window does not act on.  A 4x4 complex block on
// (QL, Rx) is the 8x8 real matrix product  D[group][out] = sum_k A[group][k] B[k][out]  over k = (Rx, QL, re/im): the
// A fragments are the thread's own registers with Rx = 0 / 1 (two chained DMMAs), the D fragment lands in the same
// two registers, and B is two doubles per thread taken from the block's matrix -- no data movement at all.  Any of
// the five register qubits can pair with QL; a block on two register qubits first exchanges one of them with QL
// (M_SWAPQL: half of the registers cross to lane ^ 2 by shuffle).
constexpr int kMmaRegBits = 5;
constexpr int kMmaWinBits = kMmaRegBits + 1;   // qubits a window can act on
constexpr int kMmaMinTileBits = 9;             // one warp
constexpr int kMmaFlagSwapOut = 1;
constexpr int kMmaDeadShift = 1;
enum : int32_t {    // WinOp codes of DMMA windows; rb / rb2 are 4-bit fields, qsel and flags 8-bit fields
    M_U2 = 16,      // dense block: rb = x (register bit paired with QL; for x != 0 the block also exchanges the roles of
                    // register bits 0 and x -- its results land in adjacent registers), rb2 = mode:
                    //   0: 4x4, matrix index bit 0 = QL, bit 1 = Rx     1: 4x4, index bit 0 = Rx, bit 1 = QL
                    //   2: 2x2 on Rx (identity on QL)                   3: 2x2 on QL (identity on Rx)
                    //   4: scalar m[0] / m[3] selected by physical bit qsel (diagonal block outside the window)
                    // flags bit kMmaFlagSwapOut (modes 0..3): the block's two output qubits trade places (matrix rows
                    // permuted when it is staged): afterwards QL holds the register qubit and register bit 0 the old QL
                    // flags bits 1..5 (kMmaDeadShift): register bits (as they sit BEFORE the block) whose qubit no gate
                    // has populated yet when the circuit starts from |0...0> -- those register pairs are zeros
    M_SWAPQL = 17,  // rb = x: exchange the roles of QL and Rx (data crosses lanes by shuffle)
    M_CX_OUT = 18,  // rb = target register bit, qsel = physical control bit (outside the window)
                    // (a CNOT with both qubits inside the window runs as an M_U2 with its constant matrix)
    M_EXPC = 20,    // expectation class: flags = flip mask over the register bits (!= 0), rb2 bit 0 = has imaginary
                    // coefficients, rb2 bit 1 = "exchange class" (two flipped bits, only the 01 <-> 10 pairs carry a
                    // coefficient: the streaming kernel evaluates half of the pairs), rb2 bit 2 = the class has Z / Y
                    // factors outside the window (unit 0 != 0), rb2 bit 3 = exchange class whose eight coefficients are
                    // equal (a bare XX + YY coupling): the streaming kernel sums the products first.
                    // Data at eterms[t]: unit 0 = Z/Y mask outside the window; units 1..8 = cA[16]
                    // (one per register pair r < r ^ flip); units 9..16 = cB[16]
    M_EXPD = 21,    // diagonal terms with Z bits outside the window: classes over register bits 0..3 (units 0..1 at
                    // eterms[t] = 16 uint16 class counts, then one unit (outside Z mask incl. register bit 4, weight)
                    // per term); qsel = physical bit of register bit 4
    M_EXPT = 22,    // diagonal terms that live on the window's register qubits: E += sum_r |psi_r|^2 D[r], D[32] at
                    // eterms[t] (16 units)
};
struct MmaWindow {  // 32 bytes
    uint8_t rpos[kMmaRegBits];      // tile position of register bit r when the window is entered
    uint8_t qlpos;                  // tile position of lane bit 1 (QL) on entry
    uint8_t gpos[3];                // tile positions of lane bits 2, 3, 4
    uint8_t wpos[3];                // tile positions of the warp-index bits (first k - 9 used)
    uint8_t rpos_out[kMmaRegBits];  // the same after the window's M_SWAPQL ops (layout written back)
    uint8_t qlpos_out;
    uint8_t flags;                  // kWinFlagReadOnly
    uint8_t dead_wbits;             // warp-index bits whose qubit nothing has populated yet when the circuit starts from
                                    // |0...0>: warps with such a bit set hold only zeros and skip the window
    uint8_t pad[4];
    int32_t op_begin, op_end;
};
static_assert(sizeof(MmaWindow) == 32, "MmaWindow layout");
// What the kernel reads: the same window with every tile position resolved to its swizzled shared-memory slot offset
// (the bank swizzle is GF(2)-linear, so a thread's slot is the XOR of the offsets of its set bits) and to its physical
// qubit (for the control / sign bits a thread reads from its own index).
struct MmaWindowDev {  // 80 bytes
    uint16_t rslot[kMmaRegBits], rslot_out[kMmaRegBits];
    uint16_t qslot, qslot_out;
    uint16_t gslot[3], wslot[3];
    uint16_t gslot_out[3], wslot_out[3];   // exit layout of the thread bits (differs from entry only in streaming passes)
    uint8_t gphys[3], wphys[3], qlphys, flags;
    int32_t op_begin, op_end;
    uint8_t rphys[kMmaRegBits];   // physical qubits of the register bits on entry (direct global loads)
    uint8_t dead_wbits;           // see MmaWindow
    // streaming passes, first window of a pass whose input holds known zeros: bits that are dead ON ENTRY -- register bits
    // (dead_r) and lane bits QL, g0, g1, g2 (dead_l bits 0..3); such registers / threads start from 0.0 instead of a load
    uint8_t dead_r, dead_l;
    uint8_t flags2;               // kWin2*
    uint8_t pad[7];
};
static_assert(sizeof(MmaWindowDev) == 80, "MmaWindowDev layout");
constexpr int kWinU4 = sizeof(MmaWindowDev) / 16;   // 16-byte units per header
constexpr int kWin2StoreAll = 1;    // entry and exit layouts differ: every thread (idle warps too) writes its registers back
constexpr int kWin2DeadEntry = 2;   // honour dead_r / dead_l / dead_wbits on entry (compact load layout)
MmaWindowDev resolve_window(const MmaWindow& w, const struct Pass& p);
// What the streaming kernel reads: byte offsets into the tile buffer (slot << 4) and the physical index bit of every
// thread bit as a mask, so that a thread's layout costs a handful of three-input logic ops.
struct StreamWindowDev {   // 112 bytes
    uint16_t rofs[kMmaRegBits], rofs_out[kMmaRegBits];
    uint16_t qofs, qofs_out;
    uint16_t gofs[3], wofs[3];
    uint16_t gofs_out[3], wofs_out[3];
    uint32_t gmask[3], wmask[3], qlmask;   // 1 << physical qubit of lane bits 2..4, of the warp bits, of QL (on entry)
    int32_t op_begin, op_end;
    uint8_t flags, flags2, dead_wbits, dead_r, dead_l, pad[3];
    uint32_t rmask[kMmaRegBits];           // 1 << physical qubit of the register bits (on entry)
};
static_assert(sizeof(StreamWindowDev) == 112, "StreamWindowDev layout");
constexpr int kSWinU4 = sizeof(StreamWindowDev) / 16;
// Streaming variant: window widx of a pass with p.stream set; sparse = the input has the known zeros of p.support_in
MmaWindowDev resolve_window_stream(const struct Pass& p, int widx, bool sparse);
StreamWindowDev stream_window_dev(const struct Pass& p, int widx, bool sparse);
// Choose the layouts of a pass for the streaming kernel (sets p.stream; false = the pass stays on tile_pass_mma_kernel).
// Gate passes: call after schedule_windows_mma and BEFORE append_expectation_windows_mma (the expectation windows then
// pick their lanes for the store layout); expectation-only passes: call after their windows exist.
bool plan_stream_layouts(struct Pass& p, int nbits);
// Consistency check of the layouts and resolved windows (every (thread, register) of every window must address a distinct
// slot, the first entry / last exit must match the TMA box order); empty string = ok
std::string validate_stream(const struct Pass& p);

struct Pass;
inline uint32_t swizzle_slot(uint32_t j);   // slot of tile index j (defined below kSwizzleVec)

// ---- streaming passes (tq_stream.cu): tiles move between HBM and shared memory by TMA -----------------------------
// The TMA engine writes a box densely in box order, 128-byte rows (8 amplitudes: physical qubits 0..2) with the hardware
// 128-byte swizzle (16-byte chunk index ^= row index & 7).  Box position bp of a tile position therefore means the
// shared-memory slot offset (1 << bp) ^ (3 <= bp <= 5 ? 1 << (bp - 3) : 0): box positions 3, 4, 5 carry the bank-swizzle
// vectors 1, 2, 4 (like positions 0, 1, 2), everything above none.  The planner picks the box order per pass so that the
// lane qubits (QL, g0, g1) of the FIRST window's entry (load layout) and of the LAST gate window's exit (store layout)
// get three different vectors: those two accesses are conflict-free, all window exchanges in between keep the
// kSwizzleVec layout.  Positions holding known zeros on input are not loaded at all (compact box).
constexpr int kStreamMaxOps = 32;    // TMA operations per tile and direction
constexpr int kStreamOpSlots = 48, kStreamWinSlots = 32;   // ops / window headers resident in shared memory (per launch)
constexpr int kStreamTileBits = 12;  // the streaming kernel's tile (two groups of 256 threads, three 64 KiB buffers per SM)
struct StreamLayout {
    int n_live = 0;               // tile positions inside the box
    uint8_t box_of[16];           // tile position -> box position, 0xff = not loaded (known zero)
    int n_dims = 1;               // dims in use; dim 0 = 16 doubles, its coordinate carries the tile's base offset
    uint8_t dim_bit[5], dim_len[5];   // dims 1..4: lowest physical bit and number of bits of the run
    int n_ops = 1;                // operations per tile: box positions beyond the four run dims are enumerated
    uint32_t op_goff[kStreamMaxOps];  // physical amplitude offset of operation i
    uint32_t box_bytes = 0;       // bytes per operation (its shared-memory destination is i * box_bytes)
    uint16_t slot(int pos) const {    // shared-memory slot offset of tile position pos
        const int bp = box_of[pos];
        return (uint16_t)((1u << bp) ^ ((bp >= 3 && bp <= 5) ? (1u << (bp - 3)) : 0u));
    }
};

struct Pass {
    uint64_t support_in = ~0ull;  // qubits some earlier gate has mixed, for a circuit started from |0...0> (tensor-core
                                  // windows put still-untouched qubits on the warp index so those warps can idle)
    bool mma = false;           // windows are MmaWindow (tensor-core kernel) instead of Window
    bool direct = false;        // expectation-only tensor-core pass whose windows all load straight from global memory
    std::vector<MmaWindow> mwindows;
    std::vector<int> local;     // physical bits of the tile, ascending; local[p] = physical bit of tile position p
    std::vector<int> nonlocal;  // remaining physical bits, ascending
    std::vector<DevOp> ops;     // tile-level ops in a valid execution order (what the windows were scheduled from)
    int lead = 0;               // number of leading positions with local[p] == p (contiguous run in memory)
    std::vector<Window> windows;  // register-window schedule of `ops` (+ expectation windows after them)
    std::vector<WinOp> wops;
    int n_gate_windows = 0;       // windows [0, n_gate_windows) change the state, the rest only read it
    std::vector<EUnit> eterms;    // data of the expectation-window ops (see W_EXPC / W_EXPD)
    // streaming kernel (plan_stream_layouts): load layouts for a dense input and for an input with the known zeros of
    // support_in, store layout; lane_vec = bank-swizzle vector per tile position of the layout the expectation windows read
    bool stream = false;
    bool sparse_differs = false;  // lin_sparse is a proper sub-box of the tile
    StreamLayout lin_dense, lin_sparse, lout;
    uint8_t lane_vec[16] = {1, 2, 4, 3, 5, 6, 7, 1, 2, 4, 3, 5, 6, 0, 0, 0};
};

struct PlanOptions {
    int tile_bits = 12;   // k: tile = 2^k amplitudes (64 KiB of complex128)
    int low_bits = 4;     // c: physical bits 0..c-1 are local in every pass (2^c * 16 B contiguous runs)
    bool trajectory = false;  // TQ_DEPOL* become per-element sampled Pauli gates (else skipped on the pure path)
    bool fuse = true;     // false: every gate is its own block (debugging / A-B comparisons)
    bool mma = true;      // pure-state passes with >= 2^9-amplitude tiles use DMMA windows
    int dead_budget = 5;  // still-empty qubits a tensor-core pass may take once the populated state spans many tiles (pack();
                          // 0 = no limit)
    bool skip_last_store = true;   // the last gate pass may leave the state unwritten when the expectation-only passes read
                                   // groups its gates do not touch (ExpPlan::last_store_needed)
    bool early_expect = true;   // Hamiltonian groups that no later gate touches may be evaluated in an earlier gate pass
                                // when that saves an expectation-only pass (attach_expectation)
    bool pack_search = false;   // TQ_PACK_SEARCH=1: besides the first-fit choice of a pass's local qubits, try a set grown qubit
                                // (pair) by qubit (pair) for the number of blocks it lets the pass execute, and keep whichever
                                // executes more -- only once every qubit is populated (generic circuits; brick chains that
                                // grow out of |0...0> keep their plans).  Off by default: measured on plans only so far.
};

struct CompiledCircuit {
    std::vector<Pass> passes;
    std::vector<MatDesc> mats;
    std::vector<MatGate> prog;
};

// Gate kinds / qubit ranges of a user circuit (what the plan_* entry points check first); false + message on error
bool validate_gates(int n, const std::vector<Gate>& gates, std::string* err);

// Pure-state plan over n qubits.  cover_masks: flip masks (physical bits) of the Hamiltonian groups the LAST pass
// should try to keep local.
CompiledCircuit plan_statevector(int n, const std::vector<Gate>& gates, const PlanOptions& opt,
                                 const std::vector<uint64_t>& cover_masks, std::string* err);

// Density-matrix plan over 2n bits: each block acts on bit q and, conjugated, on bit q + n; TQ_DEPOL* are the
// exact channels.
CompiledCircuit plan_density(int n, const std::vector<Gate>& gates, const PlanOptions& opt, std::string* err);

// Expectation-only passes that cover the flip masks `todo` (those not local in the last gate pass).
std::vector<Pass> plan_cover(int n, const std::vector<uint64_t>& todo, const PlanOptions& opt,
                             std::vector<int>* assignment);
// the local-qubit sets plan_cover would use (one per pass), without building the passes
std::vector<uint64_t> cover_sets(int n, const std::vector<uint64_t>& todo, const PlanOptions& opt,
                                 std::vector<int>* assignment);

void schedule_windows(Pass& p);
void schedule_windows_mma(Pass& p);   // needs >= kMmaMinTileBits local qubits, no density-matrix ops
// diag_pool: diagonal terms not evaluated yet (shared by the passes of a plan, consumed window by window); the final
// pass of the plan takes whatever is left.
void append_expectation_windows_mma(Pass& p, const std::vector<ExpGroupIn>& groups, std::vector<int>* leftover,
                                    std::vector<ExpTermIn>* diag_pool, bool final_pass);
// Appends read-only windows that evaluate the given Hamiltonian groups (flip masks must be local to the pass) on
// registers; groups that flip more than kRegBits qubits are returned in `leftover` (shared-memory fallback).
void append_expectation_windows(Pass& p, const std::vector<ExpGroupIn>& groups, std::vector<int>* leftover);
bool mask_is_local(const Pass& p, uint64_t mask);
// Hamiltonian groups -> passes: groups whose flips are local to the last gate pass are evaluated there, the rest in
// expectation-only passes (plan_cover) appended to `passes`; the expectation windows of every pass are appended.
// stream: choose streaming layouts (plan_stream_layouts) for the tensor-core passes first.
struct ExpPlan {
    int n_gate_passes = 0;
    std::vector<std::vector<int>> groups_of_pass;   // indices into `groups`
    std::vector<std::vector<int>> wide_of_pass;     // of those, the groups left to the shared-memory fallback
    bool last_store_needed = true;   // false: the expectation-only passes may read the INPUT of the last gate pass (no
                                     // gate of that pass touches a qubit of their groups: light cone)
    std::string err;
};
ExpPlan attach_expectation(std::vector<Pass>& passes, const std::vector<ExpGroupIn>& groups, const PlanOptions& opt, int n,
                           bool in_pass_pref, bool stream);
uint32_t mask_to_local(const Pass& p, uint64_t mask);

}  // namespace tq
