// dist.h -- multi-GPU plumbing: one context per rank (one process per GPU), NCCL over NVLink.
//
// The reference is single-process (threads + temp files) and has no collective anywhere
// (SURVEY 5 / 8e); its two shard keys -- the hash partition of a k-mer (bs/kmer_counter.h:167-178)
// and the prefix partition of a suffix (bs/part_repo.cpp:32-45) -- are the keys the exchanges
// below route by.  NCCL is bound at run time (dlopen of libnccl.so.2), so libbgx.so loads on a
// box without NCCL and single-GPU builds never touch it.
#pragma once

#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include <cuda_runtime.h>

namespace bgx {

struct Context;

struct Dist {
  int nranks = 1;
  int rank = 0;
  void* comm = nullptr;  // ncclComm_t
  // peers' device blocks mapped into this process (cudaIpcOpenMemHandle), keyed by the 64-byte handle
  std::map<std::string, void*> ipc_cache;
  // one copy stream per peer for the copy-engine exchanges
  std::vector<cudaStream_t> copy_streams;
  std::vector<cudaEvent_t> copy_done;
  cudaEvent_t copy_go = nullptr;
};

// rank 0: a fresh NCCL unique id (128 bytes) to hand to every rank out of band
void dist_get_unique_id(uint8_t id[128]);
void dist_init(Context* c, int nranks, int rank, const uint8_t id[128]);
void dist_destroy(Context* c);

// ---- collectives on device buffers, enqueued on the context's stream ---------------------------------
void dist_allreduce_sum_u64(Context* c, unsigned long long* buf, size_t n);
// every rank contributes `bytes` bytes; recv holds nranks * bytes
void dist_allgather_bytes(Context* c, const void* send, void* recv, size_t bytes);
// personalised exchange: rank r gets send[send_off[r] .. +send_cnt[r]) and writes what rank s
// sent it to recv[recv_off[s] .. +recv_cnt[s]) (units: elements of elem_bytes).  One grouped
// ncclSend/ncclRecv batch = an all-to-all over NVSwitch.
void dist_alltoallv(Context* c, const void* send, const uint64_t* send_off, const uint64_t* send_cnt, void* recv,
                    const uint64_t* recv_off, const uint64_t* recv_cnt, size_t elem_bytes);
// grouped point-to-point batch for irregular layouts
struct P2P {
  const void* send = nullptr;
  void* recv = nullptr;
  size_t bytes = 0;
  int peer = 0;
};
void dist_p2p_batch(Context* c, const std::vector<P2P>& sends, const std::vector<P2P>& recvs);

// ---- peer memory (one node: NVLink / NVSwitch) ------------------------------------------------------------
// Collective.  Every rank passes one whole block of the device arena; on return peer[r] is rank r's
// block as THIS process can address it (peer[rank] = local).  A kernel may then store straight into
// the other GPUs' memory over NVLink -- the k-mer exchange is done by the partition pass itself.
// Mappings are cached per handle (the arena hands out the same blocks run after run).  Doubles as a
// barrier: when it returns, every rank has passed the point of the call on its stream.
void dist_map_peers(Context* c, void* local, void** peer);
// the same for k blocks with one exchange: peer[j * nranks + r] = block j of rank r
void dist_map_peers_n(Context* c, void* const* local, int k, void** peer);
// every rank has passed this point on its stream (and the host has waited for it)
void dist_barrier(Context* c);
// How the two bulk exchanges of the path (k-mer words to their owners, records to theirs) move:
//   COPY  (default) the producer writes its own memory; the blocks then go to the owners' memory with
//         one copy-engine transfer per peer (cudaMemcpyAsync into the peer mapping), all peers at once
//   STORE the producing kernel stores straight into the owners' memory (fused compute + exchange).
//         Measured on B200 / NVSwitch: SM-issued 8-byte remote stores reach ~190-260 GB/s per GPU, a
//         quarter of what the copy engines move, so this form loses (profiles/r2g_*); kept for A/B.
//   NCCL  grouped ncclSend/ncclRecv (the round-1 form); kept for A/B.
enum class Exchange { COPY, STORE, NCCL };
Exchange dist_exchange_mode();   // BGX_EXCHANGE=copy|store|nccl
struct PeerCopy {
  void* dst = nullptr;      // in the peer's memory (a dist_map_peers pointer) or local
  const void* src = nullptr;
  size_t bytes = 0;
  int peer = 0;
};
// all copies at once, one stream per peer, ordered after everything queued on the context's stream;
// the context's stream continues when they are done (the data has then LEFT; dist_barrier tells when
// everybody's has arrived)
void dist_peer_copies(Context* c, const std::vector<PeerCopy>& copies);

// ---- small host-side metadata (counts, boundaries): staged through the device, synchronous -------------
// out[r * n .. (r+1) * n) = rank r's in[0..n)
void dist_allgather_host_u64(Context* c, const uint64_t* in, size_t n, uint64_t* out);
// in place: v[i] = sum over ranks of v[i]
void dist_allreduce_sum_host_u64(Context* c, uint64_t* v, size_t n);

}  // namespace bgx
