"""Stage-by-stage comparison of the smoke() step with the CPU oracle (diagnostic)."""
import sys, torch
sys.path.insert(0, '.')
import lowlightenvironmentvideoobjectdetection_b200 as vod
from oracle import vod_oracle as O
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
torch.manual_seed(0)
C, H, W, N, T, D = 64, 12, 20, 24, 3, 128
head = vod.SelsaRoIHead(
    bbox_roi_extractor=dict(type='TemporalRoIAlign', num_most_similar_points=2, num_temporal_attention_blocks=4,
                            roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=2), out_channels=C,
                            featmap_strides=[16]),
    bbox_head=dict(type='SelsaBBoxHead', num_shared_fcs=2, in_channels=C, fc_out_channels=D, num_classes=5,
                   aggregator=dict(type='SelsaAggregator', in_channels=D, num_attention_blocks=2))).to(dev)
torch.nn.init.normal_(head.bbox_head.fc_cls.weight, 0, 0.3)
torch.nn.init.normal_(head.bbox_head.fc_reg.weight, 0, 0.05)
x = torch.relu(torch.randn(1, C, H, W, generator=g))
ref_x = torch.cat([torch.relu(torch.randn(T - 1, C, H, W, generator=g)), x], 0)
def props(n):
    c = torch.rand(n, 2, generator=g) * torch.tensor([W * 16.0, H * 16.0])
    wh = torch.rand(n, 2, generator=g) * 120 + 24
    b = torch.cat([c - wh / 2, c + wh / 2], 1)
    b[:, 0::2] = b[:, 0::2].clamp(0, W * 16.0)
    b[:, 1::2] = b[:, 1::2].clamp(0, H * 16.0)
    return b
proposals = [props(N)]
ref_proposals = [props(N) for _ in range(T)]
metas = [dict(img_shape=(H * 16, W * 16, 3), scale_factor=(1.0, 1.0, 1.0, 1.0))]
rois, ref_rois = vod.bbox2roi(proposals), vod.bbox2roi(ref_proposals)
sd = {k: v.detach().cpu() for k, v in head.state_dict().items()}
ext = head.bbox_roi_extractor
rel = lambda a, b: float((a.cpu() - b).abs().max() / b.abs().max())
for kp in (False, True):
    ext.keyproj = kp
    f1 = ext((x.to(dev),), rois.to(dev), ref_feats=(ref_x.to(dev),))
    f0 = O.temporal_roi_align(x, rois, ref_x, sd['bbox_roi_extractor.embed_network.conv.weight'],
                              sd['bbox_roi_extractor.embed_network.conv.bias'], 2, 4)
    print('keyproj', kp, 'troi rel err', rel(f1, f0))
ext.keyproj = None
r1 = ext((ref_x.to(dev),), ref_rois.to(dev))
r0 = O.roi_align(ref_x, ref_rois, 7, 1 / 16, 2, True)
print('ref roi rel err', rel(r1, r0))
hp = {k[len('bbox_head.'):]: v for k, v in sd.items() if k.startswith('bbox_head.')}
cls0, reg0 = O.selsa_bbox_head(f0, r0, hp, 2, 2)
cls1, reg1 = head.bbox_head(f1, r1)
print('cls rel err', rel(cls1, cls0), 'reg rel err', rel(reg1, reg0))
d0, l0 = O.get_bboxes(rois, cls0, reg0, (H * 16, W * 16, 3), (1.0, 1.0, 1.0, 1.0), False, 0.0001, dict(type='nms', iou_threshold=0.5), 100)
dets, labels = head.simple_test((x.to(dev),), (ref_x.to(dev),), [p.to(dev) for p in proposals], [p.to(dev) for p in ref_proposals], metas, rescale=False)
d1, l1 = dets[0].cpu(), labels[0].cpu()
print(d1.shape, d0.shape, 'labels equal', torch.equal(l1, l0))
diff = (d1 - d0).abs().max(1).values
bad = (diff > 1e-2).nonzero().flatten().tolist()
print('rows differing', bad)
for i in bad[:6]:
    print(i, 'ours', d1[i].tolist(), int(l1[i]), '| oracle', d0[i].tolist(), int(l0[i]))
