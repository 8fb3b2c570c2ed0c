"""CPU: the product's DenoiseLoop (windows, per-window UNet, window accumulate, CFG, DDIM; frame sharding
and the CFG-split sharding over gloo) against the pipeline oracle and against itself on one process, with
every kernel call routed to the CPU statement of its C-ABI contract (tests/ops_contract_cpu.py)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

KW = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="linear", clip_sample=False, steps_offset=1,
          prediction_type="v_prediction", rescale_betas_zero_snr=True, timestep_spacing="trailing")


def _build(K, seed=0):
    from mikudance_b200 import synth
    from mikudance_b200.engine import UNetEngine
    from mikudance_b200.unet_3d import UNet3DConditionModel
    cfg = synth.TINY_CONFIG
    m = UNet3DConditionModel(block_out_channels=cfg["block_out_channels"],
                             cross_attention_dim=cfg["cross_attention_dim"], use_inflated_groupnorm=True,
                             use_motion_module=True, motion_module_mid_block=True, motion_module_type="Vanilla",
                             unet_use_cross_frame_attention=False, unet_use_temporal_attention=False)
    sd = synth.synthetic_state_dict(cfg, seed=seed)
    m.load_state_dict(sd)
    m = m.half().eval()
    m._engine = K.engine_on_cpu(UNetEngine, m)      # the product constructor refuses CPU models
    return cfg, m, sd


def _run_loop(m, cfg, F_, h, w, steps, ctxf, ov, pg=None):
    from mikudance_b200 import synth
    from mikudance_b200.denoise import DenoiseLoop
    from mikudance_b200.scheduler import DDIMScheduler
    lat, ctx = synth.synthetic_inputs(cfg, 2, F_, h, w, lctx=5)
    lat = lat[:1].half().contiguous()
    loop = DenoiseLoop(m, DDIMScheduler(**KW), guidance_scale=3.5, context_frames=ctxf, context_stride=1,
                       context_overlap=ov, process_group=pg, use_cuda_graph=False)
    loop.prepare(lat.clone(), ctx, steps, lambda wdw: synth.synthetic_banks(cfg, 2 * len(wdw), h, w, seed=300 + wdw[0]))
    return loop.run().float(), loop, lat, ctx


def test_denoise_loop_matches_pipeline_oracle(monkeypatch):
    import ops_contract_cpu as K
    from mikudance_b200 import synth
    from oracle.ddim_oracle import DDIMOracle
    from oracle.pipeline_oracle import denoise_loop
    K.install(monkeypatch)
    cfg, m, sd = _build(K)
    F_, h, w, steps, ctxf, ov = 6, 8, 8, 2, 4, 2          # the case of tests/test_unet_gpu.py's loop test
    got, loop, lat, ctx = _run_loop(m, cfg, F_, h, w, steps, ctxf, ov)
    assert [list(x) for x in loop.windows] == [[0, 1, 2, 3], [2, 3, 4, 5], [4, 5, 0, 1]]
    with torch.no_grad():
        want = denoise_loop({k: v.float() for k, v in sd.items()}, cfg, lat.float(), ctx.half().float(), steps, 3.5,
                            lambda wdw: synth.synthetic_banks(cfg, 2 * len(wdw), h, w, seed=300 + wdw[0]),
                            context_frames=ctxf, context_overlap=ov, scheduler=DDIMOracle(**KW))
    rel = ((got - want).norm() / want.norm()).item()
    # 2 DDIM steps x ~300 fp16-stored ops, guidance x3.5, 1x1 pixels at the deepest level (8x8 latents):
    # the GPU kernels meet 5e-3 on this case; the op-by-op CPU contract is allowed a little more
    assert rel < 8e-3, rel


def _worker(rank, world, port, q, cfg_split, F_, h, w, ctxf, ov):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["MDK_CFG_SPLIT"] = "1" if cfg_split else "0"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ops_contract_cpu as K
        from mikudance_b200 import ops
        for n in K._NAMES:
            setattr(ops, n, getattr(K, n))
        torch.set_num_threads(2)
        cfg, m, _ = _build(K)
        os.environ["MDK_CFG_SPLIT"] = "0"
        single, _, _, _ = _run_loop(m, cfg, F_, h, w, 2, ctxf, ov)
        os.environ["MDK_CFG_SPLIT"] = "1" if cfg_split else "0"
        cfg, m2, _ = _build(K)
        sharded, loop, _, _ = _run_loop(m2, cfg, F_, h, w, 2, ctxf, ov, pg=dist.group.WORLD)
        rel = ((sharded - single).norm() / single.norm()).item()
        q.put((rank, rel, loop.branch, loop.sub_world, bool(torch.isfinite(sharded).all())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,cfg_split,F_,ctxf,ov", [(2, False, 6, 4, 2), (2, True, 5, 4, 2), (4, True, 6, 4, 2)])
def test_sharded_denoise_loop_equals_single_process(world, cfg_split, F_, ctxf, ov):
    """Frame sharding (both CFG branches per rank) and the CFG split (one branch per half of the ranks)
    reproduce the single-process loop; every rank holds the same replicated latents afterwards."""
    port = 35500 + (os.getpid() % 2000) + 3 * world + int(cfg_split)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, cfg_split, F_, 8, 8, ctxf, ov)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    half = world // 2
    for rank, rel, branch, sub_world, finite in res:
        assert finite and rel < 1e-4, (rank, rel)   # float64 contract math: sharding changes only fp32 sum order
        if cfg_split:
            assert branch == rank // half and sub_world == half
        else:
            assert branch == -1 and sub_world == world


def test_plan_ranks():
    from mikudance_b200.sharding import plan_ranks
    assert plan_ranks(3, 8, True, False) == dict(branch=-1, sub_rank=3, sub_world=8)
    assert plan_ranks(3, 8, True, True) == dict(branch=0, sub_rank=3, sub_world=4)
    assert plan_ranks(5, 8, True, True) == dict(branch=1, sub_rank=1, sub_world=4)
    assert plan_ranks(1, 2, True, True) == dict(branch=1, sub_rank=0, sub_world=1)
    assert plan_ranks(1, 2, False, True)["branch"] == -1          # no guidance: nothing to split
    assert plan_ranks(1, 3, True, True)["branch"] == -1           # odd world
    assert plan_ranks(0, 1, True, True)["branch"] == -1


def test_denoise_loop_without_guidance_matches_oracle(monkeypatch):
    """guidance_scale <= 1: one branch, every image reads its bank (pipeline_mikudance.py:397, 626-630, 670-674)."""
    import ops_contract_cpu as K
    from mikudance_b200 import synth
    from mikudance_b200.denoise import DenoiseLoop
    from mikudance_b200.scheduler import DDIMScheduler
    from oracle.ddim_oracle import DDIMOracle
    from oracle.pipeline_oracle import denoise_loop
    K.install(monkeypatch)
    cfg, m, sd = _build(K)
    F_, h, w, steps = 3, 8, 8, 2
    lat, ctx = synth.synthetic_inputs(cfg, 2, F_, h, w, lctx=5)
    lat = lat[:1].half().contiguous()

    def banks(wdw):
        return synth.synthetic_banks(cfg, len(wdw), h, w, seed=500 + wdw[0])
    loop = DenoiseLoop(m, DDIMScheduler(**KW), guidance_scale=1.0, context_frames=30, use_cuda_graph=False)
    loop.prepare(lat.clone(), ctx[1:], steps, banks)
    assert loop.nb == 1 and [len(x) for x in loop.windows] == [3]
    got = loop.run().float()
    with torch.no_grad():
        want = denoise_loop({k: v.float() for k, v in sd.items()}, cfg, lat.float(), ctx.half().float(), steps, 1.0,
                            banks, context_frames=30, scheduler=DDIMOracle(**KW))
    rel = ((got - want).norm() / want.norm()).item()
    assert rel < 8e-3, rel
