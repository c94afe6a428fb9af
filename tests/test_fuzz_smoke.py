"""A few seeds of every CPU sweep of tools/fuzz_lowering.py (the full sweeps are recorded in profiles/r2_fuzz_lowering.txt):
keeps the tool alive and re-checks the lowering / host logic on inputs the fixed-seed tests do not use."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import fuzz_lowering as FZ  # noqa: E402


@pytest.mark.parametrize("mode,seeds", [("one_seed", range(7000, 7006)), ("one_ansatz_seed", range(7000, 7006)),
                                        ("one_tile_seed", range(7000, 7003)), ("one_circuit_seed", range(7000, 7008)),
                                        ("one_trajectory_seed", range(7000, 7008)), ("one_mitigation_seed", range(7000, 7004))])
def test_sweep_modes(mode, seeds):
    run = getattr(FZ, mode)
    for seed in seeds:
        assert run(seed), (mode, seed)
