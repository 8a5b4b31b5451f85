"""CPU oracle for the thermal-nerfacto per-ray hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU (torch, float32) restatement of the reference's
`implementation="torch"` algorithm for the path named in BASELINE.json:north_star.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it, and
only as the checker or the timed CPU baseline -- never from `nerfstudio_thermal_b200/` (the product
path has no CPU fallback and raises when the CUDA library is missing).

Parity pin: every function here is checked bit-for-bit (integer results) or to <=1e-6 (float
results, same op order) against golden vectors produced by importing the UNMODIFIED reference from
/root/reference in the build container (`tests/golden/make_golden.py`, fixtures in
`tests/golden/*.npz`, test `tests/test_oracle_golden.py`).  The reference's own tests pin no numeric
values on this path, so those generated vectors are the pin.

Device: every function builds its constants on the device of its inputs, so the same code runs on `cuda`
unchanged -- that is the north star's literal parity target (the reference's `implementation="torch"` ops on the
B200) and what the full-size `-m gpu` parity tests compare the kernels with (tests/test_gpu_fullsize.py).

All file:line citations are relative to /root/reference/nerfstudio/.
"""

from .hashgrid import hash_scalings, hash_corner_indices, hash_encode  # noqa: F401
from .fields import (  # noqa: F401
    mlp_forward,
    scene_contraction_linf,
    sh4_basis,
    trunc_exp,
    normalize_positions,
    density_field,
    proposal_density,
    colour_head,
)
from .sampling import (  # noqa: F401
    piecewise_spacing,
    piecewise_spacing_inv,
    initial_bins,
    pdf_resample,
    sample_weights,
    sample_positions,
    OracleSamples,
)
from .render import render_colour, render_accumulation, render_depth_median, render_depth_expected  # noqa: F401
from .model import OracleConfig, thermal_nerfacto_forward, thermal_nerfacto_losses  # noqa: F401
