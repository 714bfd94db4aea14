"""CPU: compile the device butterfly header for the host and check every
lazy-reduction variant (forward VAR 0/1/2 integer, VAR 3/4 all-FP64 incl. the shared-memory-twiddle row pass, inverse GVAR 0/1) against the
textbook merged NTT for prime sizes 30..61 bits, incl. worst-case lazy inputs."""
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))


import pytest


@pytest.mark.parametrize("frnd", [0, 1])
def test_butterfly_variants_on_host(frnd):
    """frnd = 1: the quotient rounded by cvt.rni (HEON_FP_FRND=1, two roundings) instead of the magic constant."""
    exe = os.path.join(tempfile.mkdtemp(), "host_emul")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I/usr/local/cuda/include", "-D__forceinline__=inline",
                           f"-DHEON_FP_FRND={frnd}", "-w", "-ffp-contract=off", "-o", exe, os.path.join(HERE, "host_emul.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
