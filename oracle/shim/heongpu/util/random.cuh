/* TEST INFRASTRUCTURE ONLY: stand-in for the reference's random.cuh (RNGonGPU AES DRBG +
 * OpenSSL entropy).  The context sources only initialise the generator; table
 * construction never draws from it. */
#ifndef HEONGPU_RANDOM_H
#define HEONGPU_RANDOM_H
#include <vector>
#include <cstddef>
namespace rngongpu
{
    enum class SecurityLevel { AES128, AES192, AES256 };
}
inline int RAND_bytes(unsigned char* buf, size_t num)
{
    for (size_t i = 0; i < num; ++i)
        buf[i] = (unsigned char) (i * 37 + 11);
    return 1;
}
namespace heongpu
{
    class RandomNumberGenerator
    {
      public:
        static RandomNumberGenerator& instance()
        {
            static RandomNumberGenerator r;
            return r;
        }
        void initialize(const std::vector<unsigned char>&, const std::vector<unsigned char>&,
                        const std::vector<unsigned char>&, rngongpu::SecurityLevel, bool)
        {
        }
    };
} // namespace heongpu
#endif
