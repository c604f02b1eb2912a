"""Bench-only comparison arm: the per-layer ATen op sequence the reference issues on a GPU
(norm -> argsort -> slice -> sort -> expand -> gather x2 -> cat x2; ~10-15 launches per layer),
expressed once over this repo's planner descriptors.  NOT the reference's code and NOT a product path:
it exists so the harness bench can show "one fused launch" next to "eager torch ops" on the same B200."""

import torch

from kvcompress import _planner as P


def _select(keys, plan):
    region = keys[:, :, plan.sel_lo:plan.sel_hi]
    norms = torch.norm(region, p=2, dim=-1)
    if plan.score == P.SCORE_L2_LOW:
        order = norms.argsort(dim=-1)
        picked = order[:, :, :plan.k_sel]
    elif plan.score == P.SCORE_L2_HIGH:
        order = norms.argsort(dim=-1, descending=True)
        picked = order[:, :, :plan.k_sel]
    else:  # snapkv-lite score
        imp = (norms.max(dim=-1, keepdim=True)[0] + 1e-6) - norms
        pk = plan.pool_kernel
        if pk > 1 and imp.size(-1) >= pk:
            b, h, n = imp.shape
            imp = torch.nn.functional.avg_pool1d(imp.reshape(b * h, 1, n), pk, 1, pk // 2).reshape(b, h, -1)[..., :n]
        picked = torch.topk(imp, plan.k_sel, dim=-1)[1]
    picked = torch.sort(picked, dim=-1)[0] + plan.sel_lo
    return picked


def eager_apply(kv, plans):
    out = []
    for (keys, values), plan in zip(kv, plans):
        if plan.kind == P.KEEP:
            out.append((keys, values))
            continue
        if plan.kind == P.VIEW:
            out.append((keys[:, :, -plan.view_n:], values[:, :, -plan.view_n:]))
            continue
        parts_k, parts_v = [], []
        if plan.sink:
            parts_k.append(keys[:, :, :plan.sink])
            parts_v.append(values[:, :, :plan.sink])
        if plan.k_sel:
            idx = _select(keys, plan).unsqueeze(-1).expand(-1, -1, -1, keys.size(-1))
            parts_k.append(torch.gather(keys, 2, idx))
            parts_v.append(torch.gather(values, 2, idx))
        if plan.tail:
            parts_k.append(keys[:, :, -plan.tail:])
            parts_v.append(values[:, :, -plan.tail:])
        out.append((torch.cat(parts_k, dim=2), torch.cat(parts_v, dim=2)))
    return out


def eager_fn(method):
    """fn(kv, skip_layers=..., **kwargs) with the drop-in signature, running eager torch ops."""
    from kvcompress.slab_cache import _PLANNERS, _method_defaults

    planner, names = _PLANNERS[method]

    def fn(past_key_values, skip_layers=None, **kwargs):
        kv = list(past_key_values)
        args = dict(_method_defaults(method))
        args.update(kwargs)
        if skip_layers is not None:
            args["skip_layers"] = skip_layers
        plans = planner([k.size(2) for k, _ in kv], *[args[n] for n in names], args["skip_layers"])
        return eager_apply(kv, plans)

    return fn
