import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pf = importlib.import_module('3d_mot_differentiable_pose_estimation_b200')
n = 125000
d = pf.synth.make_objects(n, 64, 64, seed=5000, device='cuda')
kinv = pf.default_kinv('cuda')
g = (torch.randn(n, device='cuda'), torch.randn(n, 9, device='cuda'), torch.randn(n, 3, device='cuda'))
def step():
    raw = pf.pose_fit_raw(d['noc'], d['depth'], d['mask'], d['bbox_xy0'], kinv)
    return pf.pose_fit_backward_raw(d['noc'], d['depth'], d['mask'], None, d['bbox_xy0'], kinv, raw.ctx, raw.status, *g)
def timeit(fn, k=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / k
print('eager %.3f ms' % timeit(step))
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
graph = torch.cuda.CUDAGraph()
with torch.cuda.stream(side):
    step(); side.synchronize()
    with torch.cuda.graph(graph, stream=side):
        out = step()
torch.cuda.current_stream().wait_stream(side)
print('graph %.3f ms' % timeit(graph.replay))
print('eager %.3f ms' % timeit(step))
