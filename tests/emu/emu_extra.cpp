// Symbols the emulated library needs beyond the kernel sources (TEST INFRASTRUCTURE ONLY).
#include "../../thor_slam_b200/csrc/ti_common.cuh"

// NCCL / peer buffers are not emulated: multi-GPU host logic is tested with gloo instead.
void ti_nccl_teardown(ti_ctx*) {}

extern "C" {
int ti_emu_marker(void) { return 1; }
int ti_nccl_unique_id(void*) { return ti::fail(nullptr, TI_ENCCL, "NCCL is not available in the CPU emulation"); }
int ti_nccl_init(ti_ctx* c, const void*, int, int) { return ti::fail(c, TI_ENCCL, "NCCL is not available in the CPU emulation"); }
int ti_gather_clouds(ti_ctx* c, const void*, void*, const uint64_t*, int) { return ti::fail(c, TI_ENCCL, "no NCCL in emulation"); }
int ti_gather_wait(ti_ctx* c, int) { return ti::fail(c, TI_ENCCL, "no NCCL in emulation"); }
int ti_gather_counts(ti_ctx* c, const uint32_t*, uint32_t*) { return ti::fail(c, TI_ENCCL, "no NCCL in emulation"); }
int ti_gather_counts_begin(ti_ctx* c, const uint32_t*) { return ti::fail(c, TI_ENCCL, "no NCCL in emulation"); }
int ti_gather_counts_finish(ti_ctx* c, uint32_t*) { return ti::fail(c, TI_ENCCL, "no NCCL in emulation"); }
int ti_gather_records(ti_ctx* c, const uint64_t*, uint64_t*, const uint32_t*, int) { return ti::fail(c, TI_ENCCL, "no NCCL in emulation"); }
int ti_exchange_fence(ti_ctx* c, uint64_t*) { return ti::fail(c, TI_ENCCL, "no exchange stream in emulation"); }
int ti_exchange_wait(ti_ctx* c, uint64_t, int) { return ti::fail(c, TI_ENCCL, "no exchange stream in emulation"); }
int ti_inbox_init(ti_ctx* c, void*) { return ti::fail(c, TI_ECUDA, "no peer memory in emulation"); }
int ti_cloud_push(ti_ctx* c, uint64_t*, const uint32_t*, uint64_t, void*, uint64_t, uint32_t) { return ti::fail(c, TI_ECUDA, "no peer memory in emulation"); }
int ti_inbox_take(ti_ctx* c, void*, uint64_t, uint32_t, uint64_t*, uint64_t, uint32_t*) { return ti::fail(c, TI_ECUDA, "no peer memory in emulation"); }
int ti_nccl_barrier(ti_ctx* c) { return ti::fail(c, TI_ENCCL, "no NCCL in emulation"); }
int ti_peer_alloc(ti_ctx* c, uint64_t, void**, void*) { return ti::fail(c, TI_ECUDA, "no peer memory in emulation"); }
int ti_peer_open(ti_ctx* c, const void*, void**) { return ti::fail(c, TI_ECUDA, "no peer memory in emulation"); }
int ti_peer_close(ti_ctx* c, void*) { return ti::fail(c, TI_ECUDA, "no peer memory in emulation"); }
int ti_peer_free(ti_ctx* c, void*) { return ti::fail(c, TI_ECUDA, "no peer memory in emulation"); }
}
