"""GPU: the attention kernel's lazy-rescale slow path (validated on hardware at the end of round 1)."""
import pytest
import torch

from tango_b200 import lib as L
from test_kernels_gpu import attn_ref, bf, rel_err, to_split

pytestmark = pytest.mark.gpu

@pytest.mark.parametrize("nsplit", [1, 2])
def test_attention_growing_scores_take_the_rescale_path(cuda, nsplit):
    """Key magnitudes ramp up along the sequence, so the running row maximum outgrows the lazy-rescale threshold (2^8)
    several times: the warps must rescale O / l in tensor memory and recompute the sub-tile against the new reference
    (random inputs never exceed the threshold, so the other attention tests only cover the fast path)."""
    B, heads, Lq, Lk = 2, 2, 200, 700
    Cc = heads * 64
    g = torch.Generator(device="cpu").manual_seed(77)
    q = torch.randn(B, Lq, Cc, generator=g)
    k = torch.randn(B, Lk, Cc, generator=g) * torch.linspace(0.5, 5.0, Lk)[None, :, None]
    v = torch.randn(B, Lk, Cc, generator=g)
    if nsplit == 1:
        qb = bf(q).reshape(B * Lq, Cc).contiguous().to(cuda)
        kvb = torch.cat([bf(k), bf(v)], dim=-1).reshape(B * Lk, 2 * Cc).contiguous().to(cuda)
        out = torch.zeros(B * Lq, Cc, device=cuda, dtype=torch.bfloat16)
        L.attention(qb, kvb, kvb, out, batch=B, heads=heads, Lq=Lq, Lk=Lk, scale=0.125, k_col0=0, v_col0=Cc)
        ref = attn_ref(bf(q).float(), bf(k).float(), bf(v).float(), heads, 0.125)
        torch.cuda.synchronize()
        assert rel_err(out.view(B, Lq, Cc).cpu(), ref) < 1e-2
    else:
        qs, ks, vs = (to_split(t.reshape(-1, Cc)).to(cuda) for t in (q, k, v))
        out = torch.zeros(B * Lq, 2 * Cc, device=cuda, dtype=torch.bfloat16)
        L.attention(qs, ks, vs, out, batch=B, heads=heads, Lq=Lq, Lk=Lk, scale=0.125, nsplit=2, q_lo_off=Cc, k_lo_off=Cc,
                    v_lo_off=Cc, split_off=Cc)
        ref = attn_ref(q.double(), k.double(), v.double(), heads, 0.125)
        torch.cuda.synchronize()
        rec = out[:, :Cc].float() + out[:, Cc:].float()
        # hi/lo products carry ~2^-16 relative error per term; scores reach ~20 nats here, so allow 5e-4 (a wrong
        # rescale would be off by O(1))
        assert rel_err(rec.view(B, Lq, Cc).cpu(), ref.float()) < 5e-4
