"""The tcgen05 attention kernels (csrc/attn_tc.cu: md_attn_fwd_tc / md_attn_bwd_tc, what md_attn_fwd / md_attn_bwd dispatch
to for head_dim 64) called directly, against the CPU contract of md_attn_fwd / md_attn_bwd: packed two-head tiles, ragged
query / key counts, multi-block keys (backward beyond 256 keys), grids larger than the SM count.  A descriptor or
protocol mistake traps through the bounded mbarrier waits instead of hanging."""
import pytest
import torch

pytestmark = [pytest.mark.gpu]

BF16 = torch.bfloat16


@pytest.mark.parametrize("B,H,Tq,Tk", [(2, 3, 128, 128), (1, 1, 128, 64), (2, 2, 256, 256), (3, 4, 64, 64), (2, 5, 256, 77),
                                       (2, 3, 100, 200), (1, 2, 64, 77), (2, 2, 1024, 77), (2, 3, 64, 64), (2, 2, 77, 77), (3, 2, 50, 33),
                                       (40, 8, 256, 256), (64, 16, 64, 77), (37, 6, 130, 16)])
def test_attn_fwd_tc_matches_contract(B, H, Tq, Tk):
    from micro_diffusion_b200.ops import CudaOps
    from oracle.emu_ops import EmuOps
    hd = 64
    hsz = H * hd
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(B * Tq, 3 * hsz + 64, generator=g).to(BF16)
    kv = torch.randn(B * Tk, 2 * hsz, generator=g).to(BF16)
    o_ref = torch.zeros(B * Tq, hsz, dtype=BF16); lse_ref = torch.zeros(B, H, Tq)
    EmuOps("cpu").attn_fwd(qkv[:, :hsz], kv[:, :hsz], kv[:, hsz:], o_ref, lse_ref, B, H, Tq, Tk, hd)
    dev = torch.device("cuda:0")
    ops = CudaOps(dev)
    ops.attn_tc = True
    qd, kd = qkv.to(dev), kv.to(dev)
    o = torch.zeros(B * Tq, hsz, dtype=BF16, device=dev); lse = torch.zeros(B, H, Tq, device=dev)
    ops.attn_fwd(qd[:, :hsz], kd[:, :hsz], kd[:, hsz:], o, lse, B, H, Tq, Tk, hd)
    torch.cuda.synchronize()
    err = (o.float().cpu() - o_ref.float()).norm() / o_ref.float().norm()
    print(f'\n[attn_fwd_tc B={B} H={H} Tq={Tq} Tk={Tk}] o rel {float(err):.3e} lse maxabs {float((lse.cpu()-lse_ref).abs().max()):.3e}')
    assert err < 2e-2, err
    assert torch.allclose(lse.cpu(), lse_ref, atol=2e-3, rtol=1e-3)


@pytest.mark.parametrize("B,H,Tq,Tk", [(1, 1, 128, 128), (2, 2, 256, 256), (3, 4, 64, 64), (2, 5, 256, 77), (2, 3, 100, 200),
                                       (1, 2, 64, 77), (1, 2, 512, 77), (2, 3, 64, 64), (2, 2, 77, 77), (3, 2, 50, 33), (20, 8, 256, 256),
                                       (64, 16, 64, 77), (40, 16, 64, 64), (37, 6, 130, 16), (2, 2, 300, 144), (1, 2, 1024, 1024),
                                       (2, 3, 300, 1000), (3, 2, 128, 513)])
def test_attn_bwd_tc_matches_contract(B, H, Tq, Tk):
    from micro_diffusion_b200.ops import CudaOps
    from oracle.emu_ops import EmuOps
    hd = 64
    hsz = H * hd
    g = torch.Generator().manual_seed(2)
    qkv = torch.randn(B * Tq, 3 * hsz + 64, generator=g).to(BF16)
    kv = torch.randn(B * Tk, 2 * hsz, generator=g).to(BF16)
    do = torch.randn(B * Tq, hsz, generator=g).to(BF16)
    emu = EmuOps("cpu")
    o_ref = torch.zeros(B * Tq, hsz, dtype=BF16); lse_ref = torch.zeros(B, H, Tq)
    emu.attn_fwd(qkv[:, :hsz], kv[:, :hsz], kv[:, hsz:], o_ref, lse_ref, B, H, Tq, Tk, hd)
    dq_ref = torch.zeros(B * Tq, hsz, dtype=BF16); dkv_ref = torch.zeros(B * Tk, 2 * hsz, dtype=BF16)
    delta = torch.zeros(B, H, Tq)
    emu.attn_bwd(do, qkv[:, :hsz], kv[:, :hsz], kv[:, hsz:], o_ref, lse_ref, delta, dq_ref, dkv_ref[:, :hsz],
                 dkv_ref[:, hsz:], B, H, Tq, Tk, hd)
    dev = torch.device("cuda:0")
    ops = CudaOps(dev)
    ops.attn_tc = True
    qd, kd, dod, od, lsed = qkv.to(dev), kv.to(dev), do.to(dev), o_ref.to(dev), lse_ref.to(dev)
    dq = torch.zeros(B * Tq, hsz, dtype=BF16, device=dev); dkv = torch.zeros(B * Tk, 2 * hsz, dtype=BF16, device=dev)
    ops.attn_bwd(dod, qd[:, :hsz], kd[:, :hsz], kd[:, hsz:], od, lsed, None, dq, dkv[:, :hsz], dkv[:, hsz:], B, H, Tq, Tk, hd)
    torch.cuda.synchronize()

    def rel(a, b):
        return float((a.float().cpu() - b.float()).norm() / b.float().norm())
    print(f'\n[attn_bwd_tc B={B} H={H} Tq={Tq} Tk={Tk}] dq {rel(dq, dq_ref):.3e} dk {rel(dkv[:, :hsz], dkv_ref[:, :hsz]):.3e} dv {rel(dkv[:, hsz:], dkv_ref[:, hsz:]):.3e}')
    assert rel(dq, dq_ref) < 2e-2 and rel(dkv, dkv_ref) < 2e-2, (rel(dq, dq_ref), rel(dkv, dkv_ref))
