// Cross-cycle dependencies of the cycle kernels (fused kernel and general-path kernels).
#pragma once
#include "osc_dev_types.h"
#include "osc_math.cuh"

namespace osc {

// Cross-cycle dependencies.  A control cycle of robot i depends on the previous cycle of robot i only (integrators, POPC and
// singularity-handler memory), and block b of the fused kernel owns the same robots in every cycle.  Instead of waiting for
// the whole previous grid (griddepcontrol.wait) -- which leaves the SMs that the last, partly filled wave does not use idle:
// 65,536 robots are 1.73 waves that then cost two -- block b of cycle c + 1 waits for the word block b of cycle c publishes
// when it is done (release / acquire at gpu scope).  A block that handed robots to the general path sets the low bit of its
// word; its successor then also waits for the general-path kernel of that cycle (general_done), whose blocks are resident
// before any block of the next fused kernel can start (both are launched with programmatic stream serialization in the same
// stream), so the wait cannot starve it.  Inputs written by other work in the stream stay ordered: a kernel that does not
// call griddepcontrol.launch_dependents releases its dependents only when it has completed, and copies are not subject to
// programmatic launch at all.
DEVI uint32_t ld_acquire_u32(const uint32_t* p) {
	uint32_t v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
DEVI void st_release_u32(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
constexpr uint32_t kEpochMask = 0x7fffffffu;

// Wait until this block's robots have finished the previous cycle.  Called by every warp right before its first access to
// task state (goals staged through shared memory, integrators, handler memory) -- the kinematics and dynamics in front of that
// point only read q, which no cycle writes -- so the round trip to the flag hides behind the first third of the kernel, and
// there is no block-wide barrier (lane 0 of each warp polls, the warp reconverges).
DEVI void wait_previous_cycle(const OscProgram& P) {
	if (!P.block_epoch) return;	 // whole-grid dependency: griddepcontrol.wait at the top of the kernel (grid_dependency_wait)
	if ((threadIdx.x & 31) == 0) {
		const uint32_t want = (P.epoch - 1u) & kEpochMask;
		uint32_t f;
		while (((f = ld_acquire_u32(P.block_epoch + blockIdx.x)) >> 1) != want) __nanosleep(64);
		if (f & 1u) {
			while ((int32_t)(ld_acquire_u32(P.general_done) - (P.epoch - 1u)) < 0) __nanosleep(256);
		}
	}
	__syncwarp();
}
// measurement aid: nanosecond timestamps of every block of the last 8 cycles (tools/pipeline_trace.py)
DEVI void stamp_block(const OscProgram& P, int which) {
	if (P.block_times && threadIdx.x == 0) {
		unsigned long long t;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
		P.block_times[((size_t)(P.epoch & 7u) * gridDim.x + blockIdx.x) * 2 + which] = t;
	}
}
DEVI void grid_dependency_wait(const OscProgram& P) {
	if (!P.block_epoch) asm volatile("griddepcontrol.wait;" ::: "memory");
}
// end of the fused kernel (every thread of the block gets here): publish the cycle number
DEVI void publish_cycle(const OscProgram& P, bool handed_over) {
	if (!P.block_epoch) return;
	const int dirty = __syncthreads_or(handed_over ? 1 : 0);
	if (threadIdx.x == 0) st_release_u32(P.block_epoch + blockIdx.x, ((P.epoch & kEpochMask) << 1) | (dirty ? 1u : 0u));
}
// end of a general-path kernel: the last block to finish clears the list counter it consumed and publishes the cycle number.
// The hint for the host (a mapped host word) is only rewritten when the count changes: the kernel cannot retire before a
// write that crosses PCIe has been acknowledged, which costs microseconds of single-cycle latency.
DEVI void publish_general_done(const OscProgram& P, int32_t count, int32_t* variant_counts = nullptr) {
	// variant_counts (split blending path, osc_blend.cuh, pipelined handles only): the four list counters to clear ([4..7]: copy of the last cycle's)
	if (!P.general_done) return;
	__syncthreads();
	if (threadIdx.x == 0) {
		bool last = true;
		if (gridDim.x > 1) {
			__threadfence();
			last = atomicAdd(&P.general_done[1], 1u) == gridDim.x - 1;
			if (last) P.general_done[1] = 0u;
		}
		if (last) {
			if (variant_counts)
				for (int k = 0; k < 4; k++) {
					variant_counts[4 + k] = variant_counts[k];	// kept for osc_debug_general_path_counts
					variant_counts[k] = 0;
				}
			P.sing_count[P.sing_parity] = 0;
			st_release_u32(P.general_done, P.epoch);
			if (P.host_seen && (uint32_t)count != P.general_done[2]) {
				P.general_done[2] = (uint32_t)count;
				*(volatile int32_t*)P.host_seen = count;
			}
		}
	}
}

}  // namespace osc
