"""Where does the end-to-end step go?  Splits bench.py's e2e leg (DUET cfg-2) into host time per API call, the upload of the
panorama features alone, and the device-side spans measured with CUDA events.   python tools/diag_e2e.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

model_kind, shape, desc = bench.workload('duet_cfg2')
model, ep = bench.build(model_kind, shape, 0, 'bf16')
dev = torch.device('cuda')
with torch.no_grad():
    d = bench.device_inputs(model_kind, model, ep, dev)
    txt, img2, loss = bench.episode_prelude(model_kind, model, d)
    DEVICE_RESIDENT = ('gmap_img_embeds', 'vp_img_embeds')
    keys = [k for k in bench.step_tensor_keys(model_kind) if k in ep and torch.is_tensor(ep[k]) and k not in DEVICE_RESIDENT]
    host = {k: ep[k].pin_memory() for k in keys}
    for k in keys:
        print('%-22s %-28s %8.1f KB' % (k, tuple(host[k].shape), host[k].numel() * host[k].element_size() / 1024))
    host_lists = {'gmap_vpids': ep['gmap_vpids'], 'vp_cand_vpids': ep['vp_cand_vpids']}
    out_host = torch.empty((shape.batch, ep['gmap_img_embeds'].shape[1]), dtype=torch.float32).pin_memory()
    model.use_cuda_graphs = True
    T = {'pano_host': 0.0, 'glue_host': 0.0, 'nav_host': 0.0, 'sync': 0.0, 'total': 0.0}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    dev_pano = dev_nav = 0.0

    def step(i, timed):
        global dev_pano, dev_nav
        if i % 6 == 0:
            model.vln_bert.drop_context()
        dd = dict(d); dd.update(host); dd.update(host_lists)
        t0 = time.perf_counter()
        ev[0].record()
        pano, _ = model('panorama', {k: dd[k] for k in bench.DUET_PANO_KEYS})
        ev[1].record()
        t1 = time.perf_counter()
        dd['vp_img_embeds'] = torch.cat([torch.zeros_like(pano[:, :1]), pano], 1)
        t2 = time.perf_counter()
        ev[2].record()
        lg = model('navigation', {**{k: dd[k] for k in bench.DUET_NAV_KEYS}, 'txt_embeds': txt, 'imagine_embeds': img2,
                                  'gmap_vpids': dd['gmap_vpids'], 'vp_cand_vpids': dd['vp_cand_vpids']})['fused_logits']
        out_host.copy_(lg, non_blocking=True)
        ev[3].record()
        t3 = time.perf_counter()
        torch.cuda.current_stream().synchronize()
        t4 = time.perf_counter()
        if timed:
            T['pano_host'] += t1 - t0; T['glue_host'] += t2 - t1; T['nav_host'] += t3 - t2; T['sync'] += t4 - t3; T['total'] += t4 - t0
            dev_pano += ev[0].elapsed_time(ev[1]); dev_nav += ev[2].elapsed_time(ev[3])

    for i in range(12):
        step(i, False)
    N = 60
    for i in range(N):
        step(i, True)
    print({k: round(v / N * 1e3, 4) for k, v in T.items()}, 'ms per step (host clock)')
    print('device spans: panorama call %.4f ms, navigation call %.4f ms' % (dev_pano / N, dev_nav / N))
    # the big upload alone
    x = host['view_img_fts']; y = torch.empty_like(x, device=dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        y.copy_(x, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print('view_img_fts upload: %.4f ms (%.1f GB/s)' % (ms, x.numel() * 4 / ms / 1e6))
