// tq_plan.h -- host-side circuit compiler of libtqsim: turns a gate list into tile passes.
//
// A "pass" streams the batch of state vectors through shared memory once: every CTA loads one tile of 2^k
// amplitudes (the amplitudes that differ only in the pass's k "local" qubits), applies the pass's gates to the
// tile and writes it back.  Gates whose mixing qubits are all local can run in the pass; diagonal action on a
// non-local qubit (RZ, Z, the control of a CNOT) only needs the tile's fixed bit and is allowed too.
// The planner packs as many gates as possible into each pass (gates that act on disjoint qubits commute, so
// a gate that does not fit is deferred together with everything that later touches its qubits).
//
// Reference behaviour being replaced: qulacs applies one gate per full-state pass
// (circuit.update_quantum_state, environments/VQAs/VQE_qulacs.py:83); see DESIGN.md section 3.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace tq {

// device-side opcodes (DevOp.op); positions are tile-local bit positions unless noted
enum : int32_t {
    OP_RX = 0,        // a = pos, t = angle source (see DevOp)
    OP_RY = 1,
    OP_RZ = 2,
    OP_RZ_NL = 3,     // a = physical bit (non-local): whole tile times e^{+-i theta/2}
    OP_CNOT = 4,      // a = control pos, b = target pos
    OP_CNOT_NL = 5,   // a = physical control bit (non-local), b = target pos: X on target if the tile's bit is set
    OP_X = 6,         // a = pos
    OP_Y = 7,         // a = pos;  flag conj: -Y (column side of a density matrix)
    OP_Z = 8,         // a = pos
    OP_Z_NL = 9,      // a = physical bit
    OP_PAULI1 = 10,   // trajectory noise slot: a = pos, t = slot (code column)
    OP_PAULI2 = 11,   // a = pos(q0), b = pos(q1), t = slot
    OP_DEPOL1_DM = 12, // a = pos(q), b = pos(q + n), fixed = p
    OP_DEPOL2_DM = 13, // a = pos(qa) | pos(qb) << 8, b = pos(qa + n) | pos(qb + n) << 8, fixed = p
};

enum : int32_t { FLAG_CONJ = 1 };  // rotation / Y acts as its complex conjugate (density-matrix column side)

struct DevOp {       // 32 bytes, read by the kernels straight from global memory
    int32_t op;
    int32_t a;
    int32_t b;
    int32_t t;       // rotations: parameter column, or -1 -> use `fixed`; PAULI*: slot
    int32_t flags;
    int32_t pad;
    double fixed;    // rotations with t == -1: theta; DEPOL*: probability
};

struct Gate {        // user-level gate (tq_set_circuit)
    int32_t kind, q0, q1, pidx;
    double fixed;
};

// ---- second level: register windows inside a tile ---------------------------------------------------------------
// Every thread of a CTA keeps 2^kRegBits amplitudes of the tile in registers: the ones that differ only in the
// window's kRegBits tile positions.  Gates that mix only window positions run on registers without touching shared
// memory; the tile is re-distributed through shared memory only when the window changes.
constexpr int kRegBits = 4;
constexpr int kMinTileBits = kRegBits;   // smaller states are padded with phantom positions
constexpr int kMaxWindowOps = 256;       // ops per window (the kernel stages one window's ops at a time)

// window-level opcodes (WinOp code); RB/CB/TB are register-bit indices inside the window
enum : int32_t {
    W_ROT_X = 0,    // rb
    W_ROT_Y = 1,    // rb
    W_ROT_Z = 2,    // rb
    W_PHASE = 3,    // RZ on a bit outside the window: qsel = physical bit
    W_CX_WW = 4,    // rb = control, rb2 = target
    W_CX_OW = 5,    // qsel = physical control bit, rb = target
    W_X = 6,        // rb
    W_Y = 7,        // rb (flag conj: -Y)
    W_Z = 8,        // rb
    W_Z_OUT = 9,    // qsel = physical bit
    W_PAULI = 10,   // rb, t = slot, rb2 = shift of the 2-bit code inside the slot byte (0 or 2)
    W_DEPOL1 = 11,  // rb = row bit, rb2 = column bit, fixed = p
    W_DEPOL2 = 12,  // the window is exactly {row a, row b, col a, col b}: rb = ra | rb << 2, rb2 = ca | cb << 2
};

struct WinOp {      // 16 bytes
    uint32_t w0;    // code | rb << 8 | rb2 << 12 | qsel << 16 | flags << 24
    int32_t t;      // rotations: parameter column or -1; W_PAULI: slot
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

// bank swizzle of the shared-memory tile: amplitude j lives in slot j ^ swizzle_fold(j >> 3), a GF(2)-linear map
// that spreads any three independent tile positions over the eight 16-byte bank groups
constexpr uint8_t kSwizzleVec[13] = {1, 2, 4, 3, 5, 6, 7, 1, 2, 4, 3, 5, 6};

struct Pass {
    std::vector<int> local;     // physical bits of the tile, ascending; local[p] = physical bit of tile position p
    std::vector<int> nonlocal;  // remaining physical bits, ascending
    std::vector<DevOp> ops;     // tile-level ops in a valid execution order (what the windows were scheduled from)
    int lead = 0;               // number of leading positions with local[p] == p (contiguous run in memory)
    std::vector<Window> windows;  // register-window schedule of `ops`
    std::vector<WinOp> wops;
};

// fills p.windows / p.wops from p.ops
void schedule_windows(Pass& p);

struct PlanOptions {
    int tile_bits = 12;   // k: tile = 2^k amplitudes (64 KiB of complex128)
    int low_bits = 4;     // c: physical bits 0..c-1 are local in every pass (2^c * 16 B contiguous runs)
    bool trajectory = false;  // TQ_DEPOL* become per-element sampled Pauli slots (else skipped on the pure path)
};

// Pure-state plan over nbits = n qubits.
// cover_masks: flip masks (physical bits) of the Hamiltonian groups the LAST pass should try to keep local.
std::vector<Pass> plan_statevector(int n, const std::vector<Gate>& gates, const PlanOptions& opt,
                                   const std::vector<uint64_t>& cover_masks, std::string* err);

// Density-matrix plan over nbits = 2n: each unitary is applied to bit q and, conjugated, to bit q + n;
// TQ_DEPOL* become the exact channels.
std::vector<Pass> plan_density(int n, const std::vector<Gate>& gates, const PlanOptions& opt, std::string* err);

// Expectation-only passes that cover the flip masks `todo` (those not local in the last gate pass).
// Returns one Pass (no ops) per tile shape; `assignment[i]` = index of the pass that evaluates todo[i].
std::vector<Pass> plan_cover(int n, const std::vector<uint64_t>& todo, const PlanOptions& opt,
                             std::vector<int>* assignment);

// true if every bit of mask is local in the pass
bool mask_is_local(const Pass& p, uint64_t mask);
// tile-local image of a physical mask (bits that are not local are dropped)
uint32_t mask_to_local(const Pass& p, uint64_t mask);

}  // namespace tq
