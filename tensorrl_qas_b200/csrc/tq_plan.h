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

struct Pass {
    std::vector<int> local;     // physical bits of the tile, ascending; local[p] = physical bit of tile position p
    std::vector<int> nonlocal;  // remaining physical bits, ascending
    std::vector<DevOp> ops;     // in execution order
    int lead = 0;               // number of leading positions with local[p] == p (contiguous run in memory)
};

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
