"""Train-step glue: lr policy restatement (CPU) and one real step on the GPU."""
import pytest
import torch

from yolo_v1_b200 import synth
from yolo_v1_b200.trainstep import (LR_ADJUST_MAP, DenseNet121Yolo, ResNet50Yolo, TrainStep, learning_rate_policy,
                                    warmming_up_policy)


def test_lr_policy_matches_train_py():
    """train.py:22-32: +1e-6 per iteration up to iteration 1000, epoch map overrides."""
    lr = 0.0
    for it in range(1, 1501):
        lr = learning_rate_policy(it, 0, lr, LR_ADJUST_MAP)
    assert lr == pytest.approx(1000 * 1e-6)
    assert warmming_up_policy(1001, 0.5) == 0.5
    assert learning_rate_policy(5, 1, 0.123, LR_ADJUST_MAP) == 0.001
    assert learning_rate_policy(5000, 75, 0.123, LR_ADJUST_MAP) == 0.0001
    assert learning_rate_policy(5000, 2, 0.123, LR_ADJUST_MAP) == 0.123


def test_head_shapes_and_strides():
    """OriginResNet.py:186-189: [N,S,S,30] as a permuted view of the NCHW head output; extra stage only for S=7."""
    n7, n14 = ResNet50Yolo(7), ResNet50Yolo(14)
    assert n7.layer5 is not None and n14.layer5 is None
    with torch.no_grad():
        y = n14.eval()(torch.randn(1, 3, 448, 448))
    assert y.shape == (1, 14, 14, 30) and y.stride() == (5880, 14, 1, 196)
    assert float(y.min()) >= 0 and float(y.max()) <= 1
    with pytest.raises(ValueError):
        ResNet50Yolo(9)
    # the reference's default backbone (train.py:57): five dense blocks for S=7, four for S=14 (OriginDenseNet.py:159-161)
    d7, d14 = DenseNet121Yolo(7), DenseNet121Yolo(14)
    assert hasattr(d7.features, "denseblock5") and not hasattr(d14.features, "denseblock5")
    with torch.no_grad():
        z = d7.eval()(torch.randn(1, 3, 448, 448))
    assert z.shape == (1, 7, 7, 30) and z.stride() == (1470, 7, 1, 49)


@pytest.mark.gpu
def test_one_bf16_train_step_updates_the_network():
    torch.manual_seed(0)
    for fuse, backbone in ((True, "resnet50"), (False, "resnet50"), (True, "densenet121")):
        ts = TrainStep(S=7, batch_size=4, device="cuda:0", fuse_head=fuse, backbone=backbone)
        ts.lr, ts.epoch = 0.0, 1          # epoch 1 -> lr 1e-3 (train.py:46-54)
        images = torch.randn(4, 3, 448, 448, device="cuda").to(memory_format=torch.channels_last)
        _, target = synth.make_loss_inputs(4, 7, seed=3, p_obj=0.1, device="cuda")
        w0 = ts.net.layer6.weight.detach().clone()
        l0 = float(ts.step(images, target))
        l1 = float(ts.step(images, target))
        assert l0 == l0 and l1 == l1 and l0 > 0          # finite
        assert not torch.equal(w0, ts.net.layer6.weight)   # the gradient reached the head through the fused loss
        assert ts.lr == 0.001


@pytest.mark.gpu
def test_train_step_equals_the_reference_step_with_the_stock_loss():
    """train.py:163-172 on a fixed seed: lr policy -> forward -> loss -> zero_grad -> backward -> SGD(momentum 0.99).
    From the SAME network and optimizer state, one step through TrainStep (fused CUDA loss, eager and graph-captured)
    and one step in the reference's order with the STOCK loss swapped in -- the unmodified `YOLOLossV1` of v1Loss.py
    on CPU (oracle/_ref, when staged) or the C oracle pinned to it, whose gradient with respect to the network output
    is handed to the same backbone -- must give the same loss and the same parameter update.  Three steps, the twin
    re-synchronised before each one: a randomly initialised ResNet-50 with batch-norm over 4 images is chaotic
    (measured: a 6e-8 difference in d loss / d pred becomes 4e-3 in the stem's weight gradient through batch-norm's
    cancelling sums, and 0.6 % in the NEXT step's loss), so free-running copies cannot be compared step for step --
    that is a property of the backbone, not of the loss."""
    import copy
    from oracle import oracle as O
    from oracle import ref_loader
    S, N = 7, 4
    ref_mod = None
    if ref_loader.available():
        RefLoss, _ = ref_loader.load_reference()
        ref_mod = RefLoss(N, S, 2, 20, 5., .5, _device='cpu')
    for graph in (False, True):
        torch.manual_seed(0)
        ts = TrainStep(S=S, batch_size=N, device="cuda:0", fuse_head=False, bf16=False, channels_last=False,
                       backbone="resnet50", graph_loss=graph)
        twin = copy.deepcopy(ts.net)                     # train mode on both: the same batch statistics, as train.py runs
        opt = torch.optim.SGD(twin.parameters(), lr=0.0, momentum=0.99)                 # train.py:84
        lr, it = 0.0, 0
        ts.start_epoch(0)                                # epoch 0: the warm-up ramp of train.py:22-25 (1e-6, 2e-6, ...)
        for step in range(3):
            twin.load_state_dict(ts.net.state_dict())
            opt.load_state_dict(copy.deepcopy(ts.opt.state_dict()))
            before = [q.detach().clone() for q in ts.net.parameters()]
            images = torch.randn(N, 3, 448, 448, generator=torch.Generator().manual_seed(10 + step)).cuda()
            _, target = synth.make_loss_inputs(N, S, seed=20 + step, p_obj=0.1)
            l_ours = float(ts.step(images, target.cuda()))
            # the reference step (train.py:157-172) with the stock loss
            it += 1
            lr = learning_rate_policy(it, 0, lr, LR_ADJUST_MAP)
            for g in opt.param_groups:
                g["lr"] = lr
            pred = twin(images)
            if ref_mod is not None:
                pc = pred.detach().cpu().contiguous().requires_grad_(True)
                with ref_loader.quiet():
                    l_ref = ref_mod(pc, target)
                    l_ref.backward()
                l_ref, g_ref = float(l_ref), pc.grad
            else:
                t5, g_np = O.loss(pred.detach().cpu().contiguous().numpy(), target.numpy(), batch_size=N)
                l_ref, g_ref = float(t5[4]), torch.from_numpy(g_np)
            opt.zero_grad()
            pred.backward(g_ref.cuda())
            opt.step()
            assert abs(l_ours - l_ref) <= 1e-5 * abs(l_ref), (graph, step, l_ours, l_ref)
            assert ts.lr == lr == pytest.approx((step + 1) * 1e-6)
            # the two UPDATES, as whole vectors (per-tensor maxima of the early layers carry batch-norm's conditioning)
            num = den = dot = n1 = 0.0
            for p0, p1, p2 in zip(before, ts.net.parameters(), twin.parameters()):
                d1, d2 = (p1.detach() - p0).double(), (p2.detach() - p0).double()
                num += float(((d1 - d2) ** 2).sum())
                den += float((d2 ** 2).sum())
                n1 += float((d1 ** 2).sum())
                dot += float((d1 * d2).sum())
            assert den > 0 and n1 > 0
            assert (num / den) ** 0.5 <= 2e-2, (graph, step, (num / den) ** 0.5)
            assert dot / (den * n1) ** 0.5 >= 1 - 1e-3, (graph, step, dot / (den * n1) ** 0.5)
