"""Where does a critic step's time go?  Captures the pieces of graphed.GraphedSteps._critic_body as separate CUDA graphs and
times their replays: generator forward, penalty branch alone, Wasserstein branch alone, both branches serial / parallel,
optimiser.  (parallel ~ max of the branches: the step is bound by a dependency chain; ~ their sum: by throughput.)
    python scripts/branch_times.py [batch] [stage]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th

from musicgan_b200 import networks
from musicgan_b200.bench_train import _build
from musicgan_b200.graphed import GraphedSteps
from musicgan_b200.networks import ops

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
stage = int(sys.argv[2]) if len(sys.argv) > 2 else 7
dev = th.device("cuda", 0)
gen, disc = _build(stage, 0, dev)
opt_g = th.optim.Adam(gen.parameters(), lr=1e-3, betas=(0.0, 0.9), capturable=True, fused=True)
opt_d = th.optim.Adam(disc.parameters(), lr=1e-3, betas=(0.0, 0.9), capturable=True, fused=True)
res = 4 * 2 ** stage
gs = GraphedSteps(gen, disc, opt_g, opt_d, batch, 32, res, 0.5)
gs.x_real.copy_(th.rand(batch, 2, res, res, device=dev) * 2 - 1)
params = list(disc.parameters())
alpha, n = gs.alpha, batch
x_fake = th.empty(batch, 2, res, res, device=dev)
keep = {}


def gen_fwd():
    z = th.randn(gs.z_shape, device=dev)
    ops.prepack(gen); ops.prepack(disc)
    with th.no_grad():
        x_fake.copy_(gen(z, alpha))


def gp_branch():
    with gs._lane(params) as lane:
        gp = disc.gradient_penalty(gs.x_real, x_fake, alpha)
        keep["gp"] = lane.merge(th.autograd.grad(gp, params, allow_unused=True))


def gp_forward_only():
    with th.no_grad():
        keep["o"] = disc(gs.x_real, alpha)


def gp_inner_only():
    keep["gp_val"] = disc.gradient_penalty(gs.x_real, x_fake, alpha).detach()


def w_branch():
    with gs._lane(params) as lane:
        out = disc(th.cat([gs.x_real, x_fake], dim=0), alpha)
        d_loss = networks.wasserstein_discriminator_loss(out[:n], out[n:])
        keep["w"] = lane.merge(th.autograd.grad(d_loss, params, allow_unused=True))


def w_forward_only():
    with th.no_grad():
        keep["o2"] = disc(th.cat([gs.x_real, x_fake], dim=0), alpha)


def serial():
    w_branch(); gp_branch()


def parallel():
    main, branch = th.cuda.current_stream(), gs._branch
    branch.wait_stream(main)
    with th.cuda.stream(branch):
        gp_branch()
    w_branch()
    main.wait_stream(branch)


def optim():
    for p in params:
        if p.grad is None:
            p.grad = th.zeros_like(p)
    opt_d.step()


def timed(name, body, reps=20):
    body(); th.cuda.synchronize()
    ws = {}
    g = th.cuda.CUDAGraph()
    with ops.capture_workspaces(ws), th.cuda.graph(g, pool=gs._gd.pool()):
        body()
    for _ in range(3):
        g.replay()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    th.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    th.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1) / reps:8.3f} ms", flush=True)
    keep[name] = (g, ws)


print(f"batch {batch} stage {stage}  lanes={os.environ.get('MG_WGRAD_LANES', '1')} depth={os.environ.get('MG_WGRAD_LANE_DEPTH', '6')}")
for name, body in (("generator forward", gen_fwd), ("critic fwd (batch B)", gp_forward_only), ("critic fwd (batch 2B)", w_forward_only),
                   ("penalty value (fwd + inner bwd)", gp_inner_only), ("penalty branch", gp_branch), ("wasserstein branch", w_branch),
                   ("both serial", serial), ("both parallel", parallel), ("optimiser", optim)):
    timed(name, body)
th.cuda.synchronize()
e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
for gname, rep in (("critic graph", gs._gd), ("generator graph", gs._gg)):
    e0.record()
    for _ in range(20):
        rep.replay()
    e1.record(); th.cuda.synchronize()
    print(f"{gname:28s} {e0.elapsed_time(e1) / 20:8.3f} ms")
