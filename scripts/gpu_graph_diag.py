"""Developer tool: where does the time of the graphed drop-in loop go (graph replay vs eager Adam step)?"""
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from types import SimpleNamespace
from skelsplat_b200 import configs, synthetic, training
from skelsplat_b200.training import optimise_frame_dropin
import skelsplat_b200.training as T
cfg = configs.H36M
seq = synthetic.make_sequence(cfg, 2, seed=100)
orig = T._graphed_iterations
def patched(gaussians, tcams, heatmaps_dense, render, opt_criterion, consistency_criterion, pipe, bg, poses_2d, cfg, data_root, accumulated_grads, iterations):
    import torch
    V = len(tcams)
    params = [gaussians.get_xyz, gaussians._scaling, gaussians._rotation, gaussians._opacity]
    static_g = [torch.zeros_like(p) for p in params[1:]]
    def body(idx):
        render_pkg = render(tcams[idx], gaussians, pipe, bg)
        l2_loss, _ = opt_criterion(render_pkg["render"], heatmaps_dense[idx], poses_2d[idx, :, :2], cfg.lambda_loss_function, reduction="mean")
        loss = l2_loss + consistency_criterion(gaussians.get_xyz, data_root, reduction="mean") * cfg.lambda_consistency
        grads = torch.autograd.grad(loss, params)
        accumulated_grads[idx].copy_(grads[0])
        for dst, g in zip(static_g, grads[1:]):
            dst.copy_(g)
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for idx in range(V): body(idx)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    graphs = []
    for idx in range(V):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g): body(idx)
        graphs.append(g)
    torch.cuda.synchronize(); print("capture 4 graphs: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
    for name, fn in (("replay", lambda i: graphs[i % V].replay()), ("eager body", lambda i: body(i % V))):
        fn(0); torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(100): fn(i)
        torch.cuda.synchronize(); print("%s: %.3f ms/iteration" % (name, (time.perf_counter() - t0) * 10))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); graphs[0].replay(); e1.record(); torch.cuda.synchronize(); print("one replay, device time %.3f ms" % e0.elapsed_time(e1))
    gaussians.get_xyz.grad = accumulated_grads.mean(0); gaussians._scaling.grad, gaussians._rotation.grad, gaussians._opacity.grad = static_g
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(100):
        gaussians.update_learning_rate(i + 1)
        gaussians.get_xyz.grad = accumulated_grads.mean(0)
        with torch.no_grad(): gaussians.optimizer.step()
    torch.cuda.synchronize(); print("lr update + eager Adam step: %.3f ms" % ((time.perf_counter() - t0) * 10))
T._graphed_iterations = patched
optimise_frame_dropin(seq.frames[0], seq.cameras, cfg, device="cuda", iterations=8, cuda_graph=True)
