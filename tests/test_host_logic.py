"""Host-side bookkeeping of the fused BatchNorm-backward reductions (no GPU): fn.BnLink hands the sums a dgrad took of a
gradient tensor to the block that receives exactly that tensor - and to nobody else."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "food101-super-resolution_b200"))


def test_bn_link_hands_the_sums_over_only_for_the_very_gradient_tensor():
    from srk import fn, ops
    link = fn.BnLink()
    link.z, link.stats, link.gamma, link.beta = (torch.zeros(1),) * 4
    grad = torch.randn(4, 4)
    red = torch.randn(9)
    link.offer(red, grad)
    assert link.take(grad) is red
    assert link.z is None and link.red is None                   # consumed: nothing is kept alive
    assert link.take(grad) is None                               # a second backward pass finds nothing and reduces itself

    # another tensor (autograd summed several gradients), or the same storage modified in place: no hand-over
    link.offer(red, grad)
    assert link.take(grad + 0.0) is None
    link.offer(red, grad)
    grad.add_(1.0)
    assert link.take(grad) is None

    # an accumulator whose consumer will not run goes back clean (zero-filled, not dirty)
    acc = ops.Acc(torch.ones(16, dtype=torch.uint8))
    acc.dirty = True
    g2 = torch.randn(3)
    link.offer(acc, g2)
    assert link.take(torch.randn(3)) is None
    assert not acc.dirty and int(acc.t.sum()) == 0
