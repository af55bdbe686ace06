"""Timing of the 2-lane conditional model (configs/lens_shape/conditional_hint_8_full.py architecture) through the FrEIA shim on
CUDA: full training step (forward both lanes, NLL, backward) vs the x-lane HINT blocks alone: python time_conditional.py [B]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import hint_b200
from FrEIA.framework import InputNode, Node, OutputNode, ReversibleGraphNet
from FrEIA.modules import HierarchicalAffineCouplingBlock, HouseholderPerm, AffineCoupling, ExternalAffineCoupling, F_fully_connected
hint_b200.set_precision("tf32")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ndim_x, ndim_y, nb, h = 20, 2, 8, 68
dev = torch.device("cuda:0")
def build(two_lane):
    y_lane = [InputNode(ndim_y, name='y')]; x_lane = [InputNode(ndim_x, name='x')]
    for i in range(nb):
        if i > 0:
            y_lane.append(Node(y_lane[-1], HouseholderPerm, {'fixed': True, 'n_reflections': ndim_y}, name=f'perm_y_{i}'))
            x_lane.append(Node(x_lane[-1], HouseholderPerm, {'fixed': True, 'n_reflections': ndim_x}, name=f'perm_x_{i}'))
        x_lane.append(Node(x_lane[-1], HierarchicalAffineCouplingBlock, {'c_internal': [h, h // 2, h // 4, h // 4]}, name=f'hac_x_{i+1}'))
        if two_lane:
            x_lane.append(Node(x_lane[-1], ExternalAffineCoupling, {'F_class': F_fully_connected, 'F_args': {'internal_size': h}},
                               conditions=y_lane[-1], name=f'ac_y_to_x_{i+1}'))
        y_lane.append(Node(y_lane[-1], AffineCoupling, {'F_class': F_fully_connected, 'F_args': {'internal_size': h // 4}}, name=f'ac_y_{i+1}'))
    y_lane.append(OutputNode(y_lane[-1], name='z_y')); x_lane.append(OutputNode(x_lane[-1], name='z_x'))
    m = ReversibleGraphNet(y_lane + x_lane, verbose=False).to(dev)
    for p in m.parameters():
        if p.requires_grad: p.data = 0.005 * torch.randn_like(p)
    return m
def step(m, x, y):
    m.zero_grad(set_to_none=True)
    z_y, z_x = m([y, x])
    J = m.log_jacobian(run_forward=False)
    loss = 0.5 * (z_y.pow(2).sum(1) + z_x.pow(2).sum(1)).mean() - J.mean()
    loss.backward()
    return loss
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
x = torch.randn(B, ndim_x, device=dev); y = torch.randn(B, ndim_y, device=dev)
for two in (True, False):
    m = build(two)
    lib = hint_b200._lib.load(); n0 = lib.hint_launch_count()
    ms = t(lambda: step(m, x, y)); n1 = lib.hint_launch_count()
    with torch.no_grad(): ms_f = t(lambda: m([y, x]))
    print(f"B={B} {'two-lane (ExternalAffineCoupling + AffineCoupling)' if two else 'x-lane HINT blocks + y-lane AffineCoupling only'}: "
          f"fwd+bwd {ms:.3f} ms, forward only {ms_f:.3f} ms, library launches per step {(n1 - n0) / 13:.0f}", flush=True)
if len(sys.argv) > 2:   # host profile of the two-lane step
    import cProfile, pstats
    m = build(True)
    for _ in range(5): step(m, x, y)
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(50): step(m, x, y)
    torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(28)
