"""oracle/ -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).

A CPU restatement (numpy; float32/float64 typed exactly as the compiled Numba kernels
type them) of the one hot path of mit-acl/mppi_numba: ``MPPI_Numba.solve()`` =
traction-map sampling -> control-noise sampling -> N x M x T unicycle rollouts with cost
accumulation -> CVaR over M -> softmax-weighted control update
(reference: mppi_numba/mppi.py:186-211, terrain.py:610-694).

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs -- and there only as the checker or the
reported CPU baseline.  ``mppi_numba_b200`` (the product) never imports it; the product
fails loudly if its CUDA library is missing.

Parity pinning: the reference ships NO tests and NO golden vectors, and
its arithmetic for this path lives partly in a third-party dependency that is not under
/root/reference: ``numba.cuda.random`` (xoroshiro128+ / splitmix64 / Box-Muller; numba is
unpinned by the reference -- README.md:66 ``pip3 install numba`` -- the image has 0.65.0).
The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build
container under Numba's CUDA simulator (``oracle/make_golden.py`` -> ``tests/golden/*.npz``,
script committed) and against known-answer vectors of numba 0.65.0's generator
(``tests/test_oracle_golden.py::test_xoroshiro_known_answers``).  The simulator is the exact-math (no fast-math) variant of the
reference; where compiled typing differs from the simulator the golden
inputs are chosen so both agree (bin values that are multiples of 1/4).
"""
