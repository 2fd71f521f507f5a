"""CPU: the BC4/BC5 restatement (oracle/restate_bc4.c) against the compiled reference and the golden vectors."""
import numpy as np

import cases
from oracle.ref import BC4, BC5


def test_restatement_matches_reference_images(ref, restated):
    for name, px, fmt in cases.scalar_cases():
        if px.dtype != np.uint8:
            continue
        assert np.array_equal(ref.encode(BC4, px, fmt), restated.bc4(px)), name
        assert np.array_equal(ref.encode(BC5, px, fmt), restated.bc5(px)), name


def test_restatement_matches_reference_blocks(ref, restated):
    blocks = cases.random_scalar_blocks(2000, seed=3)
    for b in blocks:
        assert np.array_equal(ref.alpha_block(b), restated.alpha_block(b))
