/* TEST INFRASTRUCTURE ONLY: host-only stand-in, see memorypool.cuh in this directory. */
#ifndef HEONGPU_HOST_VECTOR_H
#define HEONGPU_HOST_VECTOR_H
#include <heongpu/util/memorypool.cuh>
namespace heongpu
{
    template <typename T> class HostVector : public std::vector<T>
    {
      public:
        using std::vector<T>::vector;
    };
} // namespace heongpu
#endif
