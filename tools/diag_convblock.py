import sys, copy, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from _workloads import OdeConvBlock
from pnode_b200.convblock import ConvBlockCallbacks, recognise_convblock
for dtype in (torch.float32, torch.float64):
    func = OdeConvBlock(32, dtype=dtype).cuda()
    u = torch.randn(256, 32, 32, 32, dtype=dtype, device="cuda")
    print(dtype, "recognised:", recognise_convblock(func, u))
    probe = copy.deepcopy(func)
    cb = ConvBlockCallbacks(copy.deepcopy(func), u.shape)
    with torch.no_grad():
        ref = probe(0.3, u)
        got = cb.f(0.3, u.reshape(-1)).view_as(ref)
    print("  max abs diff", float((got - ref).abs().max()), "ref max", float(ref.abs().max()))
    print("  rm diff", float((cb.layers[0][1].running_mean - probe.bn1.running_mean).abs().max()))
