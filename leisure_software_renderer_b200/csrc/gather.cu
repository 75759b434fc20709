// gather.cu -- flag kernels of the sort-first frame assembly (shsb_frame_gather, include/shsb.h).
//
// The reference assembles nothing (one CPU, one frame).  In the sort-first multi-GPU mode every rank pushes
// its pixels straight into the ROOT GPU's assembly buffer with copy-engine peer writes over NVLink; the only thing the SMs do
// is publish / await 64-bit step counters in the root's memory, one thread each:
//   arrive[rank]  written by `rank` after its pushes of a step, awaited by the root before it consumes the step
//   released      written by the root when a step's slot may be overwritten, awaited by a rank before it reuses the slot
// Waits are BOUNDED (a rank that died must not hang the others' GPUs): on time-out the kernel raises a flag in mapped host
// memory and returns; the next C-ABI call reports it.
#include "shsb_dev.cuh"

namespace shsb
{
    namespace
    {
        __device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
        {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
            return v;
        }

        __device__ __forceinline__ unsigned long long global_timer_ns()
        {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            return t;
        }

        // monotonic publish: the flag never goes backwards (a late store of an older step must not undo a newer one)
        __global__ void gather_signal_kernel(unsigned long long* flag, unsigned long long value)
        {
            __threadfence_system(); // everything this stream wrote before (copy-engine pushes are ordered by the stream) is visible first
            atomicMax_system(flag, value);
        }

        __global__ void gather_wait_kernel(const unsigned long long* flags, uint32_t n_flags, unsigned long long value, unsigned long long timeout_ns,
                                           volatile uint32_t* timed_out)
        {
            const unsigned long long t0 = global_timer_ns();
            for (uint32_t i = 0; i < n_flags; ++i)
            {
                while (ld_acquire_sys(flags + i) < value)
                {
                    __nanosleep(256);
                    if (global_timer_ns() - t0 > timeout_ns) { *timed_out = 1u; __threadfence_system(); return; }
                }
            }
        }
    }

    void launch_gather_signal(unsigned long long* flag, unsigned long long value, cudaStream_t s, uint64_t* launches)
    {
        gather_signal_kernel<<<1, 1, 0, s>>>(flag, value);
        *launches += 1;
    }

    void launch_gather_wait(const unsigned long long* flags, uint32_t n_flags, unsigned long long value, unsigned long long timeout_ns, uint32_t* timed_out_mapped,
                            cudaStream_t s, uint64_t* launches)
    {
        gather_wait_kernel<<<1, 1, 0, s>>>(flags, n_flags, value, timeout_ns, timed_out_mapped);
        *launches += 1;
    }
}
