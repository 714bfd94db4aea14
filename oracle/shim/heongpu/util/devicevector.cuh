/* TEST INFRASTRUCTURE ONLY.  Host-only stand-in for the reference's
 * devicevector.cuh (RMM-backed): a DeviceVector that keeps its words on the
 * host, so the context tables the unmodified reference sources build can be
 * read back without a GPU.  See memorypool.cuh in this directory. */
#ifndef HEONGPU_DEVICE_VECTOR_H
#define HEONGPU_DEVICE_VECTOR_H
#include <heongpu/util/memorypool.cuh>
#include <heongpu/util/hostvector.cuh>
#include "gpufft/fft.cuh"
namespace heongpu
{
    template <typename T> class DeviceVector : public std::vector<T>
    {
      public:
        explicit DeviceVector(size_t size = 0, cudaStream_t = cudaStreamDefault) : std::vector<T>(size) {}
        explicit DeviceVector(const std::vector<T>& ref, cudaStream_t = cudaStreamDefault) : std::vector<T>(ref) {}
        template <typename A>
        explicit DeviceVector(const std::vector<T, A>& ref, cudaStream_t = cudaStreamDefault)
            : std::vector<T>(ref.begin(), ref.end())
        {
        }
    };
} // namespace heongpu
#endif
