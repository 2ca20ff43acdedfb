"""Host logic of the order-free BatchNorm running-statistic update (ops.BNSchedule): the closed form
must equal nn.BatchNorm2d's sequential r <- (1-m) r + m s for any number of calls (CPU, no kernels)."""
import pytest
import torch
import torch.nn as nn

from fusiondepth_b200 import ops


@pytest.mark.parametrize("n,momentum", [(1, 0.1), (2, 0.1), (4, 0.1), (3, 0.3)])
def test_closed_form_equals_sequential_updates(n, momentum):
    g = torch.Generator().manual_seed(n)
    bn = nn.BatchNorm2d(8, momentum=momentum)
    bn.running_mean.copy_(torch.randn(8, generator=g))
    bn.running_var.copy_(torch.rand(8, generator=g) + 0.5)
    stats = [(torch.randn(8, generator=g), torch.rand(8, generator=g) + 0.1) for _ in range(n)]
    rm, rv = bn.running_mean.clone().double(), bn.running_var.clone().double()
    for m_, v_ in stats:                                   # the reference's call-by-call update
        rm = (1 - momentum) * rm + momentum * m_.double()
        rv = (1 - momentum) * rv + momentum * v_.double()
    sched = ops.BNSchedule({bn: n})
    sched.begin_step()
    for m_, v_ in stats:                                   # what fd_bn_fwd does with stat_weight >= 0
        w = sched.next_weight(bn)
        assert w >= 0
        bn.running_mean.add_(w * m_)
        bn.running_var.add_(w * v_)
    sched.end_step()
    assert torch.allclose(bn.running_mean.double(), rm, rtol=1e-6, atol=1e-7)
    assert torch.allclose(bn.running_var.double(), rv, rtol=1e-6, atol=1e-7)
    assert int(bn.num_batches_tracked) == n


def test_schedule_rejects_unexpected_call_counts():
    bn = nn.BatchNorm2d(4)
    sched = ops.BNSchedule({bn: 2})
    sched.begin_step()
    sched.next_weight(bn)
    with pytest.raises(RuntimeError):
        sched.end_step()                                   # one call missing
    sched.begin_step()
    sched.next_weight(bn); sched.next_weight(bn)
    with pytest.raises(RuntimeError):
        sched.next_weight(bn)                              # one call too many
    other = nn.BatchNorm2d(4)
    assert sched.next_weight(other) == -1.0                # unscheduled layers keep the in-place update
