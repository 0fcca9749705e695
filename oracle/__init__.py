"""CPU oracle for the switchable-precision fake-quant linear path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped
product.  The only legal importers are ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, and there only
as the checker (or the timed CPU baseline), never as the thing shipped.  The
product path (``llm_qat_on_gpt2_b200``) never imports this package and raises
when its CUDA library is missing.

What this is: a plain numpy (float32, IEEE) restatement of the reference's
PyTorch algorithm for the path named in BASELINE.json -- calibration min/max,
min-max and log fake quantisation, the STE backward, the LoRA branch, the
switchable-precision linear, the switchable LayerNorm and the GPT-2 wrapper
that calls them.  Each function cites the reference file:line it follows
(paths relative to the upstream repo root; ``p1`` = part1_switchable_precision).

Parity pin: the reference ships no numeric golden vectors for this path
(SURVEY.md section 8c).  The oracle is therefore pinned against outputs of the
reference itself, generated in the build container by
``tests/golden/make_golden.py`` (which imports the unmodified reference from
/root/reference) and committed as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays them.

Transcendental convention: the reference's ``torch.log2`` on CPU is, to
99.987 % of inputs (measured), the correctly rounded float32 logarithm.  The
oracle *defines* log2 as ``float32(log2(float64(x)))`` (correct rounding up to
the ~2^-29 double-rounding cases), which is also what the CUDA kernels
evaluate on their exact path, so level indices are comparable bit for bit.
"""

from .quant_oracle import (  # noqa: F401
    QuantizerState,
    log2_cr,
    reduce_min_max,
    collect_statistics,
    finish_calibration,
    minmax_quantize,
    log_quantize,
    fake_quantize,
    ste_backward,
)
from .layers_oracle import (  # noqa: F401
    lora_forward,
    sp_linear_forward,
    sp_linear_backward,
    switchable_layernorm_forward,
    switchable_layernorm_backward,
)
