// Device-side completion flags for the collective-free result exchange of the data-parallel forward (SURVEY.md §8e;
// the reference is a single CPU process, main/context.py:143-147, so nothing here replaces reference code).
//
// Every rank's producing kernels store logits / class-token maps / rollout straight into rank 0's receive set over
// NVLink (vitb200_bind_outputs).  What remains of a "gather" is ordering, and it is per rank, not a barrier:
//   done[r][s]  (in rank 0's memory)  <- rank r, behind its forward into set s:  monotone step counter
//   free[s]     (in rank 0's memory)  <- rank 0, behind its reads of set s
// Rank 0's reader waits for done[r][s] of the ranks it is about to read; a writer waits for free[s] only when it
// comes back to set s a full rotation later.  No rank ever waits for a peer's forward of the SAME step, so one slow
// (power-capped) GPU no longer sets the pace of the other seven every step.
//
// Counters are monotone (step index + 1), so there is no reset and no ABA; the flags live on separate 128-byte lines.
// The signal is a system-scope release store issued by a kernel that follows the forward in stream order (kernel
// completion makes the forward's peer stores visible before the next kernel of the stream starts); the wait is a
// one-thread kernel spinning on a system-scope acquire load with back-off and a watchdog (a protocol bug traps
// instead of hanging the GPU).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace vitb200 {

__global__ void flag_signal_kernel(uint32_t* flag, uint32_t value) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

__global__ void flag_wait_kernel(const uint32_t* flag, uint32_t value, unsigned long long timeout_ns) {
  unsigned long long t0 = 0;
  uint32_t spins = 0;
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (static_cast<int32_t>(v - value) >= 0) break;   // monotone counter, wrap-safe
    __nanosleep(spins < 64 ? 64 : 512);
    if ((++spins & 0x3FFu) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > timeout_ns) __trap();
    }
  }
}

}  // namespace vitb200
