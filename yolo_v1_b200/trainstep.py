"""Train-step glue around the fused loss (BASELINE config 5).

Host-side only: the backbone stays on the stock torch / cuDNN path (it is not the product); what is restated here
is what the reference's training script does around the loss call -- the learning-rate policy
(/root/reference/train.py:22-32), SGD with momentum 0.99 (:84), the step order (:163-172) -- and the ResNet-50
YOLO head of /root/reference/backbones/OriginResNet.py:110-195 built from torchvision blocks, so that the loss
kernel can be measured inside a real bf16 DDP step.  The head's sigmoid + permute are fused into the loss
(`from_logits=True`), so the network hands the loss its raw NCHW logits.
"""
import torch
import torch.nn as nn

from .graph import GraphedLoss
from .loss import YOLOLossV1

__all__ = ["warmming_up_policy", "learning_rate_policy", "ResNet50Yolo", "DenseNet121Yolo", "TrainStep", "LR_ADJUST_MAP"]

# train.py:46-54 (epoch -> lr).  The reference file carries an unresolved merge conflict on the third key: 115 on
# its HEAD side, 100 on the other; HEAD is taken here.
LR_ADJUST_MAP = {1: 0.001, 75: 0.0001, 115: 0.00001}


def warmming_up_policy(now_iter, now_lr, stop_down_iter=1000):
    """train.py:22-25: +1e-6 per iteration for the first `stop_down_iter` iterations."""
    if now_iter <= stop_down_iter:
        now_lr += 0.000001
    return now_lr


def learning_rate_policy(now_iter, now_epoch, now_lr, lr_adjust_map, stop_down_iter=1000):
    """train.py:27-32."""
    now_lr = warmming_up_policy(now_iter, now_lr, stop_down_iter)
    if now_epoch in lr_adjust_map.keys():
        now_lr = lr_adjust_map[now_epoch]
    return now_lr


class ResNet50Yolo(nn.Module):
    """OriginResNet.py:110-195 `resnet50(S=7|14)`: torchvision ResNet-50 trunk, an extra stride-2 bottleneck stage
    (`layer5`) for S=7, 1x1 conv 2048 -> 5B+C, BatchNorm, sigmoid, permute to [N,S,S,5B+C].
    `return_logits=True` stops before the sigmoid and returns the permuted VIEW of the NCHW logits (what the fused
    loss reads in place)."""

    def __init__(self, S=7, B=2, num_classes=20, return_logits=False):
        super().__init__()
        import torchvision
        from torchvision.models.resnet import Bottleneck
        if S not in (7, 14):
            raise ValueError("S must be 7 or 14")          # OriginResNet.py:225-227
        trunk = torchvision.models.resnet50(weights=None)
        self.stem = nn.Sequential(trunk.conv1, trunk.bn1, trunk.relu, trunk.maxpool)
        self.layer1, self.layer2, self.layer3, self.layer4 = trunk.layer1, trunk.layer2, trunk.layer3, trunk.layer4
        self.layer5 = None
        if S == 7:                                           # OriginResNet.py:131-132
            down = nn.Sequential(nn.Conv2d(2048, 2048, 1, stride=2, bias=False), nn.BatchNorm2d(2048))
            self.layer5 = nn.Sequential(Bottleneck(2048, 512, stride=2, downsample=down),
                                        Bottleneck(2048, 512), Bottleneck(2048, 512))
        self.layer6 = nn.Conv2d(2048, B * 5 + num_classes, 1, bias=False)   # :133
        self.bn_end = nn.BatchNorm2d(B * 5 + num_classes)                   # :134
        self.return_logits = return_logits

    def forward(self, x):
        x = self.layer4(self.layer3(self.layer2(self.layer1(self.stem(x)))))
        if self.layer5 is not None:
            x = self.layer5(x)
        x = self.bn_end(self.layer6(x))
        if not self.return_logits:
            x = torch.sigmoid(x)                             # :188
        return x.permute(0, 2, 3, 1)                         # :189 (a view; never made contiguous)


class DenseNet121Yolo(nn.Module):
    """OriginDenseNet.py:56-164 `densenet121(S=7|14)` (the reference's default backbone, train.py:57): torchvision
    DenseNet-121 features with a fifth 16-layer dense block for S=7 (`block_config=(6,12,24,16,16)`, :159-161), ReLU,
    1x1 conv 1024 -> 5B+C, BatchNorm, sigmoid, permute (:114-129)."""

    def __init__(self, S=7, B=2, num_classes=20, return_logits=False):
        super().__init__()
        from torchvision.models.densenet import DenseNet
        if S not in (7, 14):
            raise ValueError("S must be 7 or 14")          # OriginDenseNet.py:155-157
        cfg = (6, 12, 24, 16, 16) if S == 7 else (6, 12, 24, 16)
        self.features = DenseNet(growth_rate=32, block_config=cfg, num_init_features=64).features
        self.layer6 = nn.Conv2d(1024, B * 5 + num_classes, kernel_size=1, stride=1, bias=False)   # :100
        self.bn_end = nn.BatchNorm2d(B * 5 + num_classes)                                         # :101
        self.return_logits = return_logits

    def forward(self, x):
        x = torch.relu(self.features(x))
        x = self.bn_end(self.layer6(x))
        if not self.return_logits:
            x = torch.sigmoid(x)                             # :127
        return x.permute(0, 2, 3, 1)                         # :128


class TrainStep:
    """One training iteration as train.py:155-172 runs it: lr policy, forward, loss, zero_grad, backward, step.
    bf16 autocast and DistributedDataParallel are the B200 additions (config 5)."""

    def __init__(self, S=7, B=2, C=20, batch_size=16, device="cuda", ddp=False, fuse_head=True, bf16=True,
                 channels_last=True, backbone="resnet50", graph_loss=False):
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:      # 'cuda' -> the current device, explicitly
            self.device = torch.device("cuda", torch.cuda.current_device())
        arch = {"resnet50": ResNet50Yolo, "densenet121": DenseNet121Yolo}[backbone]   # train.py:56-57
        net = arch(S, B, C, return_logits=fuse_head).to(self.device)
        if channels_last:
            net = net.to(memory_format=torch.channels_last)
        self.net = nn.parallel.DistributedDataParallel(net, device_ids=[self.device.index]) if ddp else net
        self.loss = YOLOLossV1(batch_size, S, B, C, 5., .5, from_logits=fuse_head)
        # graph_loss: the loss kernel + the all-reduce of its terms vector (job-wide logging under DDP) replayed as one
        # CUDA graph (yolo_v1_b200/graph.py) instead of eager launches -- SURVEY 8(f) row 4
        self.graphed = GraphedLoss(batch_size, S, B, C, 5., .5, from_logits=fuse_head) if graph_loss else None
        self.opt = torch.optim.SGD(self.net.parameters(), lr=0.0, momentum=0.99)    # train.py:84
        self.lr, self.iter, self.epoch, self.bf16 = 0.0, 0, 0, bf16

    def start_epoch(self, epoch):
        """train.py:148 `for epoch in range(num_epochs)`: the lr map is keyed by the epoch number (train.py:158), so
        the caller's epoch loop reports it here; epoch 0 (the default) is warm-up only, as in the reference."""
        self.epoch = int(epoch)
        return self.epoch

    def close(self):
        """Release the captured loss graph (before torch.distributed.destroy_process_group())."""
        if self.graphed is not None:
            self.graphed.release()

    def step(self, images, target):
        self.iter += 1
        self.lr = learning_rate_policy(self.iter, self.epoch, self.lr, LR_ADJUST_MAP)
        for g in self.opt.param_groups:
            g["lr"] = self.lr
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.bf16):
            pred = self.net(images)
        if self.graphed is not None:
            loss = self.graphed(pred, target)  # one graph launch: loss kernel + NCCL all-reduce of the terms
        else:
            loss = self.loss(pred, target)    # bf16 or fp32 logits, permuted NCHW view, read in place
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return loss
