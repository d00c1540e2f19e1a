// Peer "mailbox" exchange over NVLink / NVSwitch peer memory (CUDA IPC), shared by the stand-alone one-shot allreduce (comm.cu)
// and the fused tails of the INT8 pass kernels: every rank owns [2 parities][P2P_MAX_RANKS slots][P2P_MAX_DOUBLES] plus arrival
// counters and maps every peer's mailbox.  A source rank's complete contribution to one exchange bumps the destination's counter
// [parity][source] by P2P_PARTS (whoever pushes it: four CTAs of the stand-alone kernel, or one thread of a fused tail), so both
// kinds of exchange share the counters; counters only grow, the two parities alternate (see p2p_allreduce_kernel).
#pragma once
#include <cstddef>

namespace picard {

constexpr int P2P_MAX_RANKS = 8;
constexpr size_t P2P_MAX_DOUBLES = 2 * 128 * 128 + 3 * 128 + 8;  // the largest packed moment buffer the core loop exchanges (N <= 128)
constexpr int P2P_PARTS = 4;
struct P2PPeers {
  double* box[P2P_MAX_RANKS];        // box[q]: rank q's mailbox as seen from this rank (box[rank] = the local allocation)
  unsigned int* flags[P2P_MAX_RANKS];
};
struct P2PCall {                     // one exchange, as the host hands it to a kernel
  P2PPeers peers;
  int rank = 0, nranks = 1, parity = 0;
  unsigned int expect = 0;           // counter value [parity][q] reaches when rank q's contribution has arrived
};
__host__ __device__ inline size_t p2p_slot(int parity, int src) { return ((size_t)parity * P2P_MAX_RANKS + src) * P2P_MAX_DOUBLES; }

}  // namespace picard
