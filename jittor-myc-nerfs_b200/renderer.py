"""OctreeRender_trilinear_fast with the reference's signature (tensorf-myc/renderer.py:12-27).

The reference slices `rays` into `chunk`-sized pieces and synchronises the device after each one
(jt.sync_all + jt.gc).  Rays are independent, so the result does not depend on the chunk size; this
implementation keeps the argument for drop-in compatibility but launches as many rays per kernel as
the workspace budget allows and never synchronises.
"""
from __future__ import annotations

import torch


def _render_streamed(rays, tensorf, N_samples, white_bg, out_host):
    """Host-resident rays (pinned) -> host-resident rgb/depth (pinned), pipelined per workspace chunk: the upload of
    chunk i+1 and the download of chunk i-1 run on a copy stream while chunk i renders."""
    dev = tensorf.device
    S = int(N_samples) if N_samples > 0 else tensorf.nSamples
    n = rays.shape[0]
    nmax = min(tensorf.max_rays_per_launch(S), max(65536, -(-n // 4)))     # at least 4 stages when the frame is large
    st = getattr(tensorf, "_stream_state", None)
    if st is None or st["n"] != n:
        st = dict(n=n, copy=torch.cuda.Stream(device=dev), rays=torch.empty((n, 6), dtype=torch.float32, device=dev),
                  rgb=torch.empty((n, 3), dtype=torch.float32, device=dev), depth=torch.empty((n,), dtype=torch.float32, device=dev))
        tensorf._stream_state = st
    cs, main = st["copy"], torch.cuda.current_stream()
    rgb_host, depth_host = out_host
    bounds = [(s, min(n, s + nmax)) for s in range(0, n, nmax)]
    cs.wait_stream(main)
    ups = []
    with torch.cuda.stream(cs):
        for s, e in bounds:
            st["rays"][s:e].copy_(rays[s:e], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
            ups.append(ev)
    flags = tensorf._flags(white_bg)
    # the chunks alternate between the caller's stream and a side stream (own workspace each, TensorVMSplit._forward_chunks):
    # tails and launch gaps of one chunk overlap with the next chunk's march
    if getattr(tensorf, "_side_stream", None) is None:
        tensorf._side_stream = torch.cuda.Stream(device=dev)
    side = tensorf._side_stream
    side.wait_stream(main)
    for i, ((s, e), ev) in enumerate(zip(bounds, ups)):
        stream = side if (i & 1) else main
        with torch.cuda.stream(stream):
            stream.wait_event(ev)
            tensorf._forward_raw(st["rays"][s:e], None, flags, S, out=(st["rgb"][s:e], st["depth"][s:e]), ws_slot=i & 1)
            done = torch.cuda.Event()
            done.record(stream)
        with torch.cuda.stream(cs):
            cs.wait_event(done)
            rgb_host[s:e].copy_(st["rgb"][s:e], non_blocking=True)
            depth_host[s:e].copy_(st["depth"][s:e], non_blocking=True)
    main.wait_stream(side)
    main.wait_stream(cs)
    return rgb_host, depth_host


def OctreeRender_trilinear_fast(rays, tensorf, chunk=4096, N_samples=-1, ndc_ray=False, white_bg=True,
                                is_train=False, device='cuda', out_host=None):
    """`out_host=(rgb [N,3], depth [N])` (pinned host tensors, extension): with pinned host `rays` the frame is rendered
    as a copy/compute pipeline and the results land in those tensors."""
    if not torch.is_tensor(rays):
        rays = torch.as_tensor(rays, dtype=torch.float32)
    if (out_host is not None and not rays.is_cuda and rays.is_pinned() and not is_train and not ndc_ray
            and type(tensorf).__name__ != "NerfPlusPlus" and rays.dim() == 2 and rays.shape[1] == 6):
        with torch.no_grad():
            rgb_map, depth_map = _render_streamed(rays, tensorf, N_samples, white_bg, out_host)
        return rgb_map, None, depth_map, None, None
    if not rays.is_cuda:
        rays = rays.to(device, non_blocking=True)
    rays = rays.reshape(-1, rays.shape[-1])[:, :6].contiguous()
    rgb_map, depth_map = tensorf(rays, is_train=is_train, white_bg=white_bg, ndc_ray=ndc_ray, N_samples=N_samples)
    return rgb_map, None, depth_map, None, None
