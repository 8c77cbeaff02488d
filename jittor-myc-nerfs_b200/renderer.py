"""OctreeRender_trilinear_fast with the reference's signature (tensorf-myc/renderer.py:12-27).

The reference slices `rays` into `chunk`-sized pieces and synchronises the device after each one
(jt.sync_all + jt.gc).  Rays are independent, so the result does not depend on the chunk size; this
implementation keeps the argument for drop-in compatibility but launches as many rays per kernel as
the workspace budget allows and never synchronises.
"""
from __future__ import annotations

import torch

from . import _lib as L


def _render_streamed(rays, tensorf, N_samples, white_bg, out_host):
    """Host-resident rays (pinned) -> host-resident rgb/depth (pinned), pipelined per chunk: the uploads run ahead on a copy
    stream, chunk i renders (TensorVMSplit._render_bounded: two compute streams, bounded workspaces, overflow check), the
    download of chunk i follows it on the copy stream.  A large frame is cut into `stream_stages` pieces (relative sizes) so that
    the copies hide.  With `tensorf.defer_overflow_check` the call returns without reading the overflow status
    (verify_renders() before the host tensors are read): a range that has to be rendered again is uploaded again from `rays`."""
    dev = tensorf.device
    S = int(N_samples) if N_samples > 0 else tensorf.nSamples
    n = rays.shape[0]
    st = getattr(tensorf, "_stream_state", None)
    if st is None or st["n"] != n:
        st = dict(n=n, copy=torch.cuda.Stream(device=dev), rays=torch.empty((n, 6), dtype=torch.float32, device=dev),
                  rgb=torch.empty((n, 3), dtype=torch.float32, device=dev), depth=torch.empty((n,), dtype=torch.float32, device=dev))
        tensorf._stream_state = st
    live = [True]          # False once this call has returned: `launch` is then a deferred repair
    cs, main = st["copy"], torch.cuda.current_stream()
    rgb_host, depth_host = out_host
    # pipeline stages (a count of equal pieces, or relative sizes).  What is exposed is the first stage's upload, the last
    # stage's download and every download that is longer than the compute still to come; 8 equal pieces measured best
    # (173 M rays/s; (1,7,7,1): 168, (1,14,1): 163, (1,3,4,4,3,1): 173 -- profiles/r02_notes.txt P)
    weights = getattr(tensorf, "stream_stages", 8)
    weights = [1] * int(weights) if isinstance(weights, int) else list(weights)
    if n < 65536 * len(weights):
        weights = [1] * max(1, n // 65536)
    cuts = [int(round(n * sum(weights[:i]) / float(sum(weights)))) // 256 * 256 for i in range(len(weights))] + [n]
    bounds = [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    cs.wait_stream(main)
    ups = []
    with torch.cuda.stream(cs):
        for s, e in bounds:
            st["rays"][s:e].copy_(rays[s:e], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
            ups.append((s, e, ev))
    flags = tensorf._flags(white_bg) | (L.EVAL_ONLY if getattr(tensorf, "fused_composite", False) else 0)

    def launch(s, e, slot, nbytes):
        if not live[0]:
            # a deferred repair (verify_renders, device synchronised): later frames and other repairs have gone through the
            # staging buffers since -- bring the range back first
            st["rays"][s:e].copy_(rays[s:e], non_blocking=True)
        else:
            # a launch may span upload stages: wait for every stage it touches
            for s0, e0, ev in ups:
                if s0 < e and e0 > s:
                    torch.cuda.current_stream().wait_event(ev)
        tensorf._forward_raw(st["rays"][s:e], None, flags, S, out=(st["rgb"][s:e], st["depth"][s:e]), ws_slot=slot, ws_bytes=nbytes)
        return tensorf._ws if slot == 0 else tensorf._ws2

    def copy_out(s, e, done):
        with torch.cuda.stream(cs):
            cs.wait_event(done)
            rgb_host[s:e].copy_(st["rgb"][s:e], non_blocking=True)
            depth_host[s:e].copy_(st["depth"][s:e], non_blocking=True)

    tensorf._render_bounded(n, S, launch, copy_out=copy_out, ranges=bounds)
    live[0] = False
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
