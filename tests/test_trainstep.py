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
