"""SmoothQuant calibration for the HF -> FT converter: activation ranges and the smoothing transform.

Host-side torch, offline — nothing here is on the timed path.  The public functions keep the names and argument order of
the reference's ``examples/llama_quant/smoothquant.py`` so its converter script reads the same:

  ``capture_activation_range(model, tokenizer, dataset, num_samples, seq_len)``   LQ/smoothquant.py:96-144
  ``smooth_gemm(gemm_weights, act_scales, layernorm_weights, layernorm_bias, alpha, weight_scales)``   :37-67
  ``apply_smoothing(scales, gemm_weights, layernorm_weights, layernorm_bias, dtype, layernorm_1p)``    :14-34
  ``smooth_ln_fcs(ln, fcs, act_scales, alpha)``                                                         :70-93

Definitions.  For a linear layer y = x W^T with W [out, in], the smoother of input channel k is
``s_k = max|x_k|^alpha / max_j|W_jk|^(1-alpha)`` (floor 1e-5); x is divided by s (folded into the preceding norm's gamma /
beta when given) and W's column k multiplied by it, which leaves y unchanged and moves dynamic range from activations to
weights.  Several matrices that share the input (q/k/v, gate/up) share one smoother built from the largest weight.

The reference feeds lambada text through a tokenizer (512 samples x 512 tokens); this image has neither datasets nor a
network, so a calibration sample may also be a tensor / list of token ids or a dict with ``"input_ids"``.
"""
from collections import defaultdict

import torch
import torch.nn as nn

try:                                                    # GPT-2 style [in, out] projections are hooked too
    from transformers.pytorch_utils import Conv1D as _Conv1D
except Exception:                                       # pragma: no cover
    _Conv1D = ()

_FLOOR = 1e-5


def _as_list(x):
    return x if isinstance(x, list) else [x]


def _column_ranges(weights):
    """largest |w| per input channel over every row of every matrix in ``weights`` ([out, in] each)"""
    per_matrix = [w.abs().amax(dim=0) for w in weights]
    return per_matrix[0] if len(per_matrix) == 1 else torch.stack(per_matrix).amax(dim=0)


@torch.no_grad()
def apply_smoothing(scales, gemm_weights, layernorm_weights=None, layernorm_bias=None, dtype=torch.float32,
                    layernorm_1p=False):
    """In place: gamma / beta of the preceding norm divided by ``scales``, every weight's input columns multiplied by it.
    (``dtype`` is accepted for interface parity: as in the reference the tensors keep their own dtype.)"""
    del dtype
    for t in (layernorm_weights, layernorm_bias):
        if t is not None:
            assert t.numel() == scales.numel()
            t.div_(scales)
    if layernorm_1p:                                    # norms stored as (gamma - 1)
        layernorm_weights += (1 / scales) - 1
    row = scales.reshape(1, -1)
    for w in _as_list(gemm_weights):
        w.mul_(row)


@torch.no_grad()
def smooth_gemm(gemm_weights, act_scales, layernorm_weights=None, layernorm_bias=None, alpha=0.5, weight_scales=None):
    """Compute the smoother of the shared input of ``gemm_weights`` from the calibrated ``act_scales`` and apply it in
    place; returns it (fp64, on the weights' device).  As in the reference the weight ranges are used in their own dtype
    and only the quotient is floored."""
    ws = _as_list(gemm_weights)
    for w in ws:
        assert w.shape[1] == act_scales.numel(), "weights are [out, in]"
    if weight_scales is None:
        weight_scales = _column_ranges(ws)
    act = act_scales.to(ws[0].device).to(torch.float64)
    scales = (act.pow(alpha) / weight_scales.pow(1 - alpha)).clamp(min=_FLOOR)
    apply_smoothing(scales, ws, layernorm_weights, layernorm_bias, ws[0].dtype)
    return scales


@torch.no_grad()
def smooth_ln_fcs(ln, fcs, act_scales, alpha=0.5):
    """Module flavour: smooth the ``nn.Linear`` layers ``fcs`` that all read the output of norm ``ln`` (RMSNorm: no bias)."""
    fcs = _as_list(fcs)
    for fc in fcs:
        assert isinstance(fc, nn.Linear) and fc.in_features == act_scales.numel() == ln.weight.numel()
    w0 = fcs[0].weight
    act = act_scales.to(device=w0.device, dtype=w0.dtype)
    wmax = _column_ranges([fc.weight for fc in fcs]).clamp(min=_FLOOR)
    scales = (act.pow(alpha) / wmax.pow(1 - alpha)).clamp(min=_FLOOR).to(w0.dtype)
    ln.weight.div_(scales)
    if getattr(ln, "bias", None) is not None:
        ln.bias.div_(scales)
    for fc in fcs:
        fc.weight.mul_(scales.reshape(1, -1))
    return scales


def _token_ids(sample, tokenizer, seq_len, device):
    if isinstance(sample, dict):
        if "input_ids" in sample:
            sample = sample["input_ids"]
        else:
            if tokenizer is None:
                raise ValueError("text samples need a tokenizer")
            return tokenizer(sample["text"], return_tensors="pt", max_length=seq_len, truncation=True).input_ids.to(device)
    ids = torch.as_tensor(sample, dtype=torch.long)
    return (ids[None] if ids.dim() == 1 else ids)[:, :seq_len].to(device)


@torch.no_grad()
def capture_activation_range(model, tokenizer, dataset, num_samples=512, seq_len=512):
    """Run ``num_samples`` prompts through ``model`` with a forward hook on every linear layer and return
    ``{module name: {"x": max |input| per channel, "y": max |output| per channel, "w": max |weight| over dim 0}}``.
    x / y statistics are fp32; "w" is taken once, clipped below at 1e-8, in the weight's dtype — over dim 0 of
    ``module.weight``, i.e. per INPUT channel for ``nn.Linear`` ([out, in]) and per output column for GPT-2's Conv1D; the
    converter replaces it with per-output-column ranges where it needs them."""
    model.eval()
    device = next(model.parameters()).device
    ranges = defaultdict(lambda: {"x": None, "y": None, "w": None})

    def fold(entry, key, t):
        cur = t.detach().reshape(-1, t.shape[-1]).abs().amax(dim=0).float()
        entry[key] = cur if entry[key] is None else torch.maximum(entry[key], cur)

    def make_hook(name):
        def hook(module, inputs, output):
            entry = ranges[name]
            fold(entry, "x", inputs[0] if isinstance(inputs, tuple) else inputs)
            fold(entry, "y", output)
            if entry["w"] is None:
                entry["w"] = module.weight.abs().clip(1e-8, None).amax(dim=0)
        return hook

    linear_types = (nn.Linear,) + ((_Conv1D,) if _Conv1D else ())
    handles = [m.register_forward_hook(make_hook(n)) for n, m in model.named_modules() if isinstance(m, linear_types)]
    try:
        for i in range(min(num_samples, len(dataset))):
            model(_token_ids(dataset[i], tokenizer, seq_len, device))
    finally:
        for h in handles:
            h.remove()
    return ranges
