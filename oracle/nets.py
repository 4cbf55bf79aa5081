"""oracle/nets.py -- TEST INFRASTRUCTURE ONLY.

PyTorch fp32 CPU restatement of every network on GLASS's inference path, with
parameter names equal to the detectron2 / reference ``state_dict`` names
 so a released ``.pth`` and the oracle's seeded weights load on both
sides.  detectron2 pieces are [d2-recall] restatements of v0.6; GLASS pieces cite the
reference file:line they follow and are pinned by tests/golden (tools/make_golden.py).
"""
import math
from typing import Dict, List

import torch
import torch.nn.functional as F
from torch import nn


# ----------------------------------------------------------------------------- d2 layers
class D2Conv2d(nn.Conv2d):
    """d2 layers/wrappers.py Conv2d: conv -> norm -> activation."""

    def __init__(self, *args, norm: bool = False, relu: bool = False, **kw):
        super().__init__(*args, **kw)
        self.norm = nn.BatchNorm2d(self.out_channels, eps=1e-5) if norm else None
        self.relu = relu

    def forward(self, x):
        x = super().forward(x)
        if self.norm is not None:
            x = self.norm(x)
        if self.relu:
            x = F.relu(x)
        return x


def c2_msra_fill(m, gen):
    # fvcore.nn.weight_init.c2_msra_fill: kaiming_normal_(mode="fan_out", nonlinearity="relu")
    fan_out = m.weight.shape[0] * m.weight[0][0].numel()
    std = math.sqrt(2.0 / fan_out)
    with torch.no_grad():
        m.weight.normal_(0, std, generator=gen)
        if m.bias is not None:
            m.bias.zero_()


def c2_xavier_fill(m, gen):
    # fvcore c2_xavier_fill: kaiming_uniform_(a=1) -> bound = sqrt(3 / fan_in)
    fan_in = m.weight[0].numel()
    bound = math.sqrt(3.0 / fan_in)
    with torch.no_grad():
        m.weight.uniform_(-bound, bound, generator=gen)
        if m.bias is not None:
            m.bias.zero_()


def normal_fill(m, std, gen):
    with torch.no_grad():
        m.weight.normal_(0, std, generator=gen)
        if getattr(m, "bias", None) is not None:
            m.bias.zero_()


def torch_default_fill(m, gen):
    # nn.Conv2d / nn.Linear default: kaiming_uniform_(a=sqrt(5)) -> bound = 1/sqrt(fan_in)
    fan_in = m.weight[0].numel()
    bound = 1.0 / math.sqrt(fan_in)
    with torch.no_grad():
        m.weight.uniform_(-bound, bound, generator=gen)
        if getattr(m, "bias", None) is not None:
            m.bias.uniform_(-bound, bound, generator=gen)


# ------------------------------------------------------------------- d2 ResNet-50 (A.2)
class BottleneckBlock(nn.Module):
    def __init__(self, cin, cout, bottleneck, stride):
        super().__init__()
        self.shortcut = D2Conv2d(cin, cout, 1, stride=stride, bias=False, norm=True) if cin != cout else None
        # STRIDE_IN_1X1 = True (d2 default; glass_finetune_totaltext.yaml:139)
        self.conv1 = D2Conv2d(cin, bottleneck, 1, stride=stride, bias=False, norm=True)
        self.conv2 = D2Conv2d(bottleneck, bottleneck, 3, stride=1, padding=1, bias=False, norm=True)
        self.conv3 = D2Conv2d(bottleneck, cout, 1, bias=False, norm=True)

    def forward(self, x):
        out = F.relu(self.conv1(x))
        out = F.relu(self.conv2(out))
        out = self.conv3(out)
        sc = self.shortcut(x) if self.shortcut is not None else x
        return F.relu(out + sc)


class BasicStem(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = D2Conv2d(3, 64, 7, stride=2, padding=3, bias=False, norm=True)

    def forward(self, x):
        return F.max_pool2d(F.relu(self.conv1(x)), kernel_size=3, stride=2, padding=1)


class ResNet50(nn.Module):
    """configs/glass_pretrain.yaml:41-50 -> d2 build_resnet_backbone, DEPTH 50."""

    def __init__(self):
        super().__init__()
        self.stem = BasicStem()
        cin = 64
        for i, (n, stage) in enumerate(zip([3, 4, 6, 3], ["res2", "res3", "res4", "res5"])):
            bott, cout = 64 * 2 ** i, 256 * 2 ** i
            blocks = []
            for b in range(n):
                blocks.append(BottleneckBlock(cin, cout, bott, stride=2 if (b == 0 and i > 0) else 1))
                cin = cout
            setattr(self, stage, nn.Sequential(*blocks))

    def forward(self, x) -> Dict[str, torch.Tensor]:
        out = {}
        x = self.stem(x)
        for stage in ["res2", "res3", "res4", "res5"]:
            x = getattr(self, stage)(x)
            out[stage] = x
        return out


class ResNetFPN(nn.Module):
    """d2 FPN + LastLevelMaxPool (glass_pretrain.yaml:51-54), A.3."""

    def __init__(self):
        super().__init__()
        self.bottom_up = ResNet50()
        for k, cin in zip([2, 3, 4, 5], [256, 512, 1024, 2048]):
            setattr(self, f"fpn_lateral{k}", D2Conv2d(cin, 256, 1, bias=False, norm=True))
            setattr(self, f"fpn_output{k}", D2Conv2d(256, 256, 3, padding=1, bias=False, norm=True))

    def forward(self, x) -> Dict[str, torch.Tensor]:
        c = self.bottom_up(x)
        prev = self.fpn_lateral5(c["res5"])
        out = {"p5": self.fpn_output5(prev)}
        for k in [4, 3, 2]:
            top_down = F.interpolate(prev, scale_factor=2.0, mode="nearest")
            prev = getattr(self, f"fpn_lateral{k}")(c[f"res{k}"]) + top_down
            out[f"p{k}"] = getattr(self, f"fpn_output{k}")(prev)
        out["p6"] = F.max_pool2d(out["p5"], kernel_size=1, stride=2, padding=0)
        out.update(c)
        return out


# -------------------------------------------------------------------- RPN head (A.4)
class RPNHead(nn.Module):
    def __init__(self, num_anchors=12, box_dim=5):
        super().__init__()
        self.conv = D2Conv2d(256, 256, 3, padding=1, relu=True)
        self.objectness_logits = nn.Conv2d(256, num_anchors, 1)
        self.anchor_deltas = nn.Conv2d(256, num_anchors * box_dim, 1)

    def forward(self, feats: List[torch.Tensor]):
        logits, deltas = [], []
        for x in feats:
            t = self.conv(x)
            logits.append(self.objectness_logits(t))
            deltas.append(self.anchor_deltas(t))
        return logits, deltas


# ------------------------------------------------------------------ box head (A.7)
class FastRCNNConvFCHead(nn.Module):
    def __init__(self, cin=256 * 7 * 7, dim=2048):
        super().__init__()
        self.fc1 = nn.Linear(cin, dim)
        self.fc2 = nn.Linear(dim, dim)

    def forward(self, x):
        x = torch.flatten(x, start_dim=1)
        return F.relu(self.fc2(F.relu(self.fc1(x))))


class RotatedFastRCNNOutputLayers(nn.Module):
    """glass/modeling/roi_heads/rotated_fast_rcnn.py:536-550 (layers), :587-599 (forward)."""

    def __init__(self, dim=2048, num_classes=1):
        super().__init__()
        self.cls_score = nn.Linear(dim, num_classes + 1)
        self.bbox_pred = nn.Linear(dim, num_classes * 5)
        self.orientation_pred = nn.Linear(dim, 4)

    def forward(self, x):
        return self.cls_score(x), self.bbox_pred(x), self.orientation_pred(x)


# ---------------------------------------------------- GLASS: P2P3Fusion / local CNN / fusion
class P2P3Fusion(nn.Module):
    """glass/modeling/fusion/fusion_modules.py:250-286."""

    def __init__(self, c=256):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 1, bias=False)
        self.conv2 = nn.Conv2d(c, c, 1, bias=False)

    def forward(self, p2, p3):
        return F.interpolate(self.conv2(p3), scale_factor=2.0, mode="nearest") + self.conv1(p2)


class LocalBasicBlock(nn.Module):
    """glass/modeling/fusion/local_feature_extraction.py:290-323."""

    def __init__(self, cin, planes):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = None
        if cin != planes:
            self.downsample = nn.Sequential(nn.Conv2d(cin, planes, 1, bias=False), nn.BatchNorm2d(planes))

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        res = x if self.downsample is None else self.downsample(x)
        return F.relu(out + res)


class LocalResNet(nn.Module):
    """glass/modeling/fusion/local_feature_extraction.py:95-188, layers [1,2,5,3]."""

    def __init__(self, cin=3, cout=256):
        super().__init__()
        blk = [cout // 4, cout // 2, cout, cout]
        self.conv0_1 = nn.Conv2d(cin, cout // 16, 3, padding=1, bias=False)
        self.bn0_1 = nn.BatchNorm2d(cout // 16)
        self.conv0_2 = nn.Conv2d(cout // 16, cout // 8, 3, padding=1, bias=False)
        self.bn0_2 = nn.BatchNorm2d(cout // 8)
        inpl = cout // 8

        def make(planes, n):
            nonlocal inpl
            layers = []
            for _ in range(n):
                layers.append(LocalBasicBlock(inpl, planes))
                inpl = planes
            return nn.Sequential(*layers)

        self.layer1 = make(blk[0], 1)
        self.conv1 = nn.Conv2d(blk[0], blk[0], 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(blk[0])
        self.layer2 = make(blk[1], 2)
        self.conv2 = nn.Conv2d(blk[1], blk[1], 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(blk[1])
        self.layer3 = make(blk[2], 5)
        self.conv3 = nn.Conv2d(blk[2], blk[2], 3, padding=1, bias=False)
        self.bn3 = nn.BatchNorm2d(blk[2])
        self.layer4 = make(blk[3], 3)
        self.conv4_1 = nn.Conv2d(blk[3], blk[3], 2, stride=(2, 1), padding=(0, 0), bias=False)
        self.bn4_1 = nn.BatchNorm2d(blk[3])

    def forward(self, x):
        x = F.relu(self.bn0_1(self.conv0_1(x)))
        x = F.relu(self.bn0_2(self.conv0_2(x)))
        if x.shape[0] > 0:
            x = F.max_pool2d(x, 2, 2, 0)
        x = F.relu(self.bn1(self.conv1(self.layer1(x))))
        if x.shape[0] > 0:
            x = F.max_pool2d(x, 2, 2, 0)
        x = F.relu(self.bn2(self.conv2(self.layer2(x))))
        if x.shape[0] > 0:
            x = F.max_pool2d(x, kernel_size=2, stride=(2, 1), padding=(0, 1))
        x = F.relu(self.bn3(self.conv3(self.layer3(x))))
        x = self.layer4(x)
        return F.relu(self.bn4_1(self.conv4_1(x)))


class ResNetFeatureExtractor(nn.Module):
    """glass/modeling/fusion/local_feature_extraction.py:22-29."""

    def __init__(self, cin=3, cout=256):
        super().__init__()
        self.ConvNet = LocalResNet(cin, cout)

    def forward(self, x):
        return self.ConvNet(x)


class MultiAspectGCAttention(nn.Module):
    """glass/modeling/fusion/fusion_modules.py:22-157 (pooling 'att', fusion 'channel_add')."""

    def __init__(self, inplanes=512, ratio=0.5, headers=8, outplane=256):
        super().__init__()
        self.headers = headers
        self.inplanes = inplanes
        self.planes = int(inplanes * ratio)
        self.single = inplanes // headers
        order = torch.zeros(inplanes, dtype=torch.long)
        order[0::2] = torch.arange(inplanes)[: inplanes // 2]
        order[1::2] = torch.arange(inplanes)[inplanes // 2:]
        self.order = order  # plain attribute, not a buffer (fusion_modules.py:50-53)
        self.out = nn.Conv2d(inplanes, outplane, 3, padding=1)
        self.conv_mask = nn.Conv2d(self.single, 1, 1)
        self.channel_add_conv = nn.Sequential(
            nn.Conv2d(inplanes, self.planes, 1), nn.LayerNorm([self.planes, 1, 1]),
            nn.ReLU(inplace=True), nn.Conv2d(self.planes, inplanes, 1))

    def spatial_pool(self, x):
        b, c, h, w = x.shape
        xh = x.reshape(b * self.headers, self.single, h, w)
        mask = self.conv_mask(xh).view(b * self.headers, 1, h * w)
        mask = F.softmax(mask, dim=2).unsqueeze(-1)                       # [B*h,1,HW,1]
        ctx = torch.matmul(xh.reshape(b * self.headers, 1, self.single, h * w), mask)
        return ctx.view(b, c, 1, 1)

    def forward(self, x):
        x = x[:, self.order, ...]
        ctx = self.spatial_pool(x)
        return self.out(x + self.channel_add_conv(ctx))


# ------------------------------------------------------------------ GLASS: recognizer
class CNN_V1_1(nn.Module):
    """glass/modeling/recognition/recognizer_backbone.py:34-81."""

    def __init__(self, c=256):
        super().__init__()
        self.conv1 = D2Conv2d(c, c, (2, 1), stride=(2, 1), padding=0, bias=False, norm=True, relu=True)
        self.conv2 = D2Conv2d(c, c, 3, stride=1, padding=1, bias=False, norm=True, relu=True)

    def forward(self, x):
        x1 = self.conv1(x)
        return self.conv2(x1) + x1


class BiLSTM(nn.Module):
    """glass/modeling/recognition/recognizer_encoder.py:123-144."""

    def __init__(self, cin, hidden, cout):
        super().__init__()
        self.rnn = nn.LSTM(cin, hidden, bidirectional=True, batch_first=True)
        self.linear = nn.Linear(hidden * 2, cout)

    def forward(self, x):
        return self.linear(self.rnn(x)[0])


class BiLSTMBlockV2(nn.Module):
    """glass/modeling/recognition/recognizer_encoder.py:100-120."""

    def __init__(self, c=256, layers=2):
        super().__init__()
        self.bilsm_stack = nn.Sequential(*[BiLSTM(c, c, c) for _ in range(layers)])

    def forward(self, feats):
        return self.bilsm_stack(feats.mean(dim=2).transpose(1, 2).contiguous())


class AttentionUnit(nn.Module):
    def __init__(self, s, x, a):
        super().__init__()
        self.sEmbed = nn.Linear(s, a)
        self.xEmbed = nn.Linear(x, a)
        self.wEmbed = nn.Linear(a, 1)


class DecoderUnit(nn.Module):
    """glass/modeling/recognition/prediction_aster.py:269-302 (+ AttentionUnit :225-266)."""

    def __init__(self, s_dim, x_dim, y_dim, att_dim):
        super().__init__()
        self.attention_unit = AttentionUnit(s_dim, x_dim, att_dim)
        self.tgt_embedding = nn.Embedding(y_dim, att_dim)
        self.gru = nn.GRU(input_size=x_dim + att_dim, hidden_size=s_dim, batch_first=True)
        self.fc = nn.Linear(s_dim, y_dim)
        self.temperature = nn.Parameter(torch.ones(1), requires_grad=False)

    def step(self, x, x_proj, s_prev, y_prev):
        au = self.attention_unit
        s_proj = au.sEmbed(s_prev.squeeze(0)).unsqueeze(1)                # [b,1,att]
        v = au.wEmbed(torch.tanh(s_proj + x_proj)).squeeze(-1)            # [b,T]
        alpha = F.softmax(v, dim=1)
        context = torch.bmm(alpha.unsqueeze(1), x).squeeze(1)
        y_proj = self.tgt_embedding(y_prev.long())
        out, state = self.gru(torch.cat([y_proj, context], 1).unsqueeze(1), s_prev)
        out = self.fc(out.squeeze(1)) * self.temperature
        return out, state, alpha


class AttentionRecognitionHead(nn.Module):
    """glass/modeling/recognition/prediction_aster.py:14-99 (inference `sample` only)."""

    def __init__(self, num_classes, in_planes, s_dim, att_dim, max_len):
        super().__init__()
        self.num_classes = num_classes
        self.s_dim = s_dim
        self.decoder = DecoderUnit(s_dim, in_planes, num_classes, att_dim)

    def sample(self, x, lengths, eos, taps=None):
        b = x.shape[0]
        state = torch.zeros(1, b, self.s_dim)
        dones = torch.zeros(b)
        outputs_ = torch.zeros(b, lengths, self.num_classes)
        logits_ = torch.zeros(b, lengths, self.num_classes)
        alphas_ = torch.zeros(b, lengths, x.shape[1])
        # xProj is step-invariant; the reference recomputes it each step (:250) -- same values.
        x_proj = self.decoder.attention_unit.xEmbed(x)
        predicted = torch.zeros((b,))
        steps = 0
        for i in range(lengths):
            y_prev = torch.zeros((b,)) if i == 0 else predicted
            output, state, alpha = self.decoder.step(x, x_proj, state, y_prev)
            prob = F.softmax(output[:, : self.num_classes], dim=1)
            _, predicted = prob.max(1)
            outputs_[:, i] = prob
            logits_[:, i] = output
            alphas_[:, i] = alpha
            steps = i + 1
            dones += (predicted == eos).float()
            if dones.min() != 0:
                break
        if taps is not None:
            taps["decoder_logits"] = logits_
            taps["decoder_alpha"] = alphas_
            taps["decoder_steps"] = steps
        return outputs_


class ASTER_V2(nn.Module):
    """glass/modeling/recognition/recognizer_decoder.py:65-93."""

    def __init__(self, num_classes=97, max_word_len=26, c=256):
        super().__init__()
        self.max_word_len = max_word_len
        self.recognizer = AttentionRecognitionHead(num_classes, c, c, c, max_word_len)

    def forward(self, feats, taps=None):
        return self.recognizer.sample(feats.contiguous(), self.max_word_len, 0, taps=taps)


class RecognizerRCNNHeadV3(nn.Module):
    """glass/modeling/recognition/recognizer_head_v2.py:291-345, inference :150-163."""

    def __init__(self, c=256, num_classes=97, max_word_len=26):
        super().__init__()
        self.backbone = CNN_V1_1(c)
        self.encoder = BiLSTMBlockV2(c, 2)
        self.decoder = ASTER_V2(num_classes, max_word_len, c)

    def forward(self, x, taps=None):
        f = self.backbone(x)
        e = self.encoder(f)
        if taps is not None:
            taps["recog_cnn"] = f
            taps["encoder_out"] = e
        return self.decoder(e, taps=taps)
