"""Stage-by-stage forward comparison of ImageResNetTrainer against the oracle (debug aid)."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from oracle import dense_oracle as DO
from tests.test_backbone_gpu import build
from tests.test_train_gpu import _bf16_sd, rel
from dhd_b200 import train as T
from dhd_b200.train_backbone import ImageResNetTrainer

for bn in ('frozen', 'batch'):
    net, neck = build('bf16')
    net = net.cpu()
    net.load_state_dict(_bf16_sd(net.state_dict()))
    img = DO.seeded_tensor((2, 3, 128, 192), 21).bfloat16().float()
    sd = net.state_dict()
    DO.BN_TRAIN = bn == 'batch'
    with torch.no_grad():
        stem = F.relu(DO._bn(sd, 'bn1', F.conv2d(img, sd['conv1.weight'], stride=2, padding=3)))
        pool = F.max_pool2d(stem, 3, stride=2, padding=1)
        want = DO.image_resnet_forward(sd, img, 50, (0, 1, 2, 3))
    DO.BN_TRAIN = False
    net = net.cuda()
    T.set_bn_mode(bn)
    tr = ImageResNetTrainer(net)
    T.set_bn_mode('frozen')
    tr.out_indices = (0, 1, 2, 3)
    outs = tr.forward(img.cuda())
    col, s_out = tr.saved_stem
    print(bn, 'stem', rel(s_out.float().cpu(), stem))
    x0 = tr.saved[0][0]
    print(bn, 'pool', rel(x0.float().cpu(), pool))
    k = 0
    for li, blocks in enumerate(tr.layers):
        for bi in range(len(blocks)):
            k += 1
        print(bn, 'layer', li, rel(outs[li].float().cpu(), want[li]), 'out absmax', float(want[li].abs().max()))
    # first block in detail
    with torch.no_grad():
        DO.BN_TRAIN = bn == 'batch'
        x = pool
        p = 'layer1.0.'
        y1 = F.relu(DO._bn(sd, p + 'bn1', F.conv2d(x, sd[p + 'conv1.weight'])))
        y2 = F.relu(DO._bn(sd, p + 'bn2', F.conv2d(y1, sd[p + 'conv2.weight'], padding=1)))
        y3 = DO._bn(sd, p + 'bn3', F.conv2d(y2, sd[p + 'conv3.weight']))
        idn = DO._bn(sd, p + 'downsample.1', F.conv2d(x, sd[p + 'downsample.0.weight']))
        o = F.relu(y3 + idn)
        DO.BN_TRAIN = False
    xs, t1, t2, out = tr.saved[0]
    print(bn, 'l1.0 t1', rel(t1.float().cpu(), y1), 't2', rel(t2.float().cpu(), y2), 'out', rel(out.float().cpu(), o))
    idb = tr._buf.get(('id_0_0', 2, x0.H, x0.W, 256))
    print(bn, 'l1.0 idn', rel(idb.float().cpu(), idn))
    xs, t1, t2, out = tr.saved[3]
    with torch.no_grad():
        DO.BN_TRAIN = bn == 'batch'
        x = want[0]
        p = 'layer2.0.'
        y1 = F.relu(DO._bn(sd, p + 'bn1', F.conv2d(x, sd[p + 'conv1.weight'])))
        y2 = F.relu(DO._bn(sd, p + 'bn2', F.conv2d(y1, sd[p + 'conv2.weight'], padding=1, stride=2)))
        idn = DO._bn(sd, p + 'downsample.1', F.conv2d(x, sd[p + 'downsample.0.weight'], stride=2))
        DO.BN_TRAIN = False
    idb = tr._buf.get(('id_1_0', 2, t2.H, t2.W, 512))
    print(bn, 'l2.0 t1', rel(t1.float().cpu(), y1), 't2', rel(t2.float().cpu(), y2), 'idn', rel(idb.float().cpu(), idn))
