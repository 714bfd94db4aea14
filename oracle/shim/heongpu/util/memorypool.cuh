/* TEST INFRASTRUCTURE ONLY.  Host-only stand-in for the reference's
 * src/include/heongpu/util/memorypool.cuh (which needs RMM, absent from this
 * image).  It lets the reference's OWN context sources (bfv/context.cu,
 * ckks/context.cu) compile UNMODIFIED for oracle/_ref/libref_ctx.so, so the
 * tables they build can be compared word for word with the product's.  Only
 * the few names those two files use are declared. */
#ifndef HEONGPU_MEMORYPOOL_H
#define HEONGPU_MEMORYPOOL_H
#include <cuda_runtime.h>
#include <memory>
#include <mutex>
#include <optional>
#include <vector>
#include "gpuntt/common/common.cuh"
#include "gpuntt/common/nttparameters.cuh"
#include <heongpu/kernel/defines.h>
#include <heongpu/util/util.cuh>
namespace heongpu
{
    struct MemoryPoolConfig
    {
        std::optional<float> initial_device_fraction, max_device_fraction;
        std::optional<size_t> initial_device_bytes, max_device_bytes;
        std::optional<float> initial_host_fraction, max_host_fraction;
        std::optional<size_t> initial_host_bytes, max_host_bytes;
        bool use_memory_pool = true;
        static MemoryPoolConfig Defaults() { return MemoryPoolConfig{}; }
    };
    class MemoryPool
    {
      public:
        static MemoryPool& instance()
        {
            static MemoryPool p;
            return p;
        }
        void initialize() {}
        void initialize(const MemoryPoolConfig&) {}
        void use_memory_pool(bool) {}
    };
    template <typename T> using rmm_pinned_allocator = std::allocator<T>;
} // namespace heongpu
#endif
