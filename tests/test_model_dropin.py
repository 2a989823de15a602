"""Model-level drop-in check (SURVEY.md §8 a-7 / §8c "model-level oracle"): the networks that call
CubePad — the cubic ResNet stem + Bottleneck (model/resnet_cubic.py:66-117) and the ConvLSTM cell
(model/clstm.py:25-83) — are rebuilt here with the reference's layer order, once around
cp360_b200.CubePad (one libcp360 launch per call site) and once around the torch slice/flip/cat port
of the reference module (oracle.ref_port.CubePadPort, test infrastructure). Same weights, same input,
same cuDNN convolutions: CubePad is pure data movement, so the activations must be BIT-IDENTICAL, and
the gradients (training path, train_temporal.py:105-107,167-170) agree to fp32 summation order.
The convolutions themselves stay torch/cuDNN (not the product, BASELINE.json north_star).
"""
import pytest
import torch
import torch.nn as nn

import cp360_b200
from oracle import ref_port

pytestmark = pytest.mark.gpu


class PortPad(nn.Module):
    def __init__(self, p):
        super().__init__()
        self.port = ref_port.CubePadPort(p)

    def forward(self, x):
        return self.port(x)


class Bottleneck(nn.Module):
    """resnet_cubic.py:66-107 (expansion 4, CubePad(1) in front of the 3x3 convolution)."""

    def __init__(self, pad, inplanes, planes, stride=1):
        super().__init__()
        self.pad = pad(1)
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=0, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.down = None
        if stride != 1 or inplanes != planes * 4:
            self.down = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False),
                                      nn.BatchNorm2d(planes * 4))

    def forward(self, x):
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(self.pad(out))))
        out = self.bn3(self.conv3(out))
        res = x if self.down is None else self.down(x)
        return self.relu(out + res)


class CubicStem(nn.Module):
    """resnet_cubic.py:109-170: CubePad(3) + 7x7/2 conv, CubePad(1) + 3x3/2 max-pool, then blocks."""

    def __init__(self, pad):
        super().__init__()
        self.pad3, self.pad1 = pad(3), pad(1)
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=0, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=0)
        self.layer1 = nn.Sequential(Bottleneck(pad, 64, 64), Bottleneck(pad, 256, 64))
        self.layer2 = nn.Sequential(Bottleneck(pad, 256, 128, stride=2))

    def forward(self, x):
        x = self.relu(self.bn1(self.conv1(self.pad3(x))))
        x = self.maxpool(self.pad1(x))
        return self.layer2(self.layer1(x))


class ConvLSTMCell(nn.Module):
    """clstm.py:25-83: cat(input, hidden) -> pad -> Conv1 -> ReLU -> pad -> Conv2 -> ReLU -> pad -> Gates."""

    def __init__(self, pad, input_size, hidden_size):
        super().__init__()
        self.hidden_size = hidden_size
        self.Conv1 = nn.Conv2d(input_size + hidden_size, 4 * hidden_size, 3, padding=0)
        self.Conv2 = nn.Conv2d(4 * hidden_size, 4 * hidden_size, 3, padding=0)
        self.Gates = nn.Conv2d(4 * hidden_size, 4 * hidden_size, 3, padding=0)
        self.pad = pad(1)

    def forward(self, x, state):
        h, c = state
        out = torch.relu(self.Conv1(self.pad(torch.cat((x, h), 1))))
        out = torch.relu(self.Conv2(self.pad(out)))
        gates = self.Gates(self.pad(out))
        i, r, o, g = gates.chunk(4, 1)
        c = torch.sigmoid(r) * c + torch.sigmoid(i) * torch.tanh(g)
        return torch.sigmoid(o) * torch.tanh(c), c


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch.device("cuda", 0)


@pytest.fixture(autouse=True)
def deterministic_cudnn():
    old = (torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark,
           torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    (torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark,
     torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32) = old


def _pair(make, dev):
    torch.manual_seed(7)
    ours = make(cp360_b200.CubePad).to(dev)
    port = make(PortPad).to(dev)
    port.load_state_dict(ours.state_dict())     # CubePad has no parameters or buffers (cube_pad.py:23-31)
    return ours, port


def test_cubic_resnet_stem_bit_identical(dev):
    ours, port = _pair(CubicStem, dev)
    ours.eval(), port.eval()
    assert not [k for k in ours.state_dict() if "pad" in k]
    x = torch.randn(12, 3, 64, 64, device=dev)            # two cubes
    before = cp360_b200._lib.launch_count()
    with torch.no_grad():
        a, b = ours(x), port(x)
    # CubePad(3), CubePad(1) and one CubePad(1) per Bottleneck: 5 launches for the whole batch
    assert cp360_b200._lib.launch_count() - before == 5
    assert tuple(a.shape) == (12, 512, 8, 8)
    assert torch.equal(a, b)


def test_cubic_resnet_stem_training_step(dev):
    """Forward in train mode (batch statistics) bit-identical; input and weight gradients agree."""
    ours, port = _pair(CubicStem, dev)
    ours.train(), port.train()
    x = torch.randn(6, 3, 48, 48, device=dev)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = ours(xa), port(xb)
    assert torch.equal(ya, yb)
    g = torch.randn_like(ya)
    ya.backward(g)
    yb.backward(g)
    scale = float(xb.grad.abs().max().item())
    assert float((xa.grad - xb.grad).abs().max().item()) <= 1e-4 * max(scale, 1.0)
    for (n, pa), (_, pb) in zip(ours.named_parameters(), port.named_parameters()):
        s = max(float(pb.grad.abs().max().item()), 1.0)
        assert float((pa.grad - pb.grad).abs().max().item()) <= 2e-4 * s, n


@pytest.mark.parametrize("feat,hidden,hw,cubes", [(1000, 1000, 7, 1), (64, 32, 8, 2)])
def test_convlstm_cell_sequence(dev, feat, hidden, hw, cubes):
    """Reference-exact ConvLSTM sites ([6,2000,7,7], [6,4000,7,7] x2 per step, clstm.py:57-64) over a
    short sequence: hidden and cell states bit-identical at every step."""
    ours, port = _pair(lambda pad: ConvLSTMCell(pad, feat, hidden), dev)
    n = 6 * cubes
    torch.manual_seed(3)
    ha = hb = torch.zeros(n, hidden, hw, hw, device=dev)
    ca = cb = torch.zeros(n, hidden, hw, hw, device=dev)
    with torch.no_grad():
        for _ in range(3):
            x = torch.randn(n, feat, hw, hw, device=dev)
            ha, ca = ours(x, (ha, ca))
            hb, cb = port(x, (hb, cb))
            assert torch.equal(ha, hb) and torch.equal(ca, cb)


def test_convlstm_cat_fusion_matches(dev):
    """cubepad_cat (cat + CubePad without the concatenated tensor, clstm.py:57-58) feeds Conv1 the same bits."""
    torch.manual_seed(5)
    x = torch.randn(6, 48, 7, 7, device=dev)
    h = torch.randn(6, 16, 7, 7, device=dev)
    want = ref_port.CubePadPort(1)(torch.cat((x, h), 1))
    assert torch.equal(cp360_b200.cubepad_cat([x, h], 1), want)


def test_saliency_training_head(dev):
    """train_temporal.py:105-107: hidden state -> to_equi_nn -> channel max -> loss.backward(), through
    the fused differentiable max and through the un-fused pair of ops."""
    w, C = 7, 50
    c2e = cp360_b200.Cube2Equi(w)
    torch.manual_seed(11)
    h = torch.randn(6, C, w, w, device=dev)
    target = torch.rand(1, 2 * w, 4 * w, device=dev)
    ha, hb = h.clone().requires_grad_(True), h.clone().requires_grad_(True)
    la = ((c2e.to_equi_max(cp360_b200.CubePad(1)(ha)[:, :, 1:-1, 1:-1]) - target) ** 2).mean()
    lb = ((c2e.to_equi_nn(ref_port.CubePadPort(1)(hb)[:, :, 1:-1, 1:-1]).max(1)[0] - target) ** 2).mean()
    assert torch.equal(la, lb)
    la.backward()
    lb.backward()
    torch.testing.assert_close(ha.grad, hb.grad, rtol=0, atol=1e-6)
