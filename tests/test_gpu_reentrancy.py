"""The C ABI is re-entrant across streams and host threads: one DjPack is shared, every
stream works in its own context (per-shape tables, error word, workspaces), every meshing thread in its own DjMcWork."""
import threading

import numpy as np
import pytest
import torch

from oracle import synth
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp16x2", "bf16", "fp32"])
def test_two_streams_share_one_pack(precision):
    dec = util.make_decoder(0, "trained_norm", precision=precision)
    codes = [util.trained_code(i).cuda() for i in range(4)]
    pts = synth.make_points(300_000, 11).cuda()
    serial = [dec.query(c, pts, use_net=1).clone() for c in codes]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(4)]
    for rep in range(3):
        outs = [None] * 4
        for i, st in enumerate(streams):  # enqueue all four before any finishes: the kernels overlap / interleave
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                outs[i] = dec.query(codes[i], pts, use_net=1)
        torch.cuda.synchronize()
        for i in range(4):
            assert torch.equal(outs[i], serial[i]), (rep, i)


def test_two_host_threads_mesh_with_one_decoder():
    from deepjoin_b200.core import reconstruct as R
    dec = util.make_decoder(0, "trained_norm", precision="fp16x2")
    codes = [util.trained_code(i) for i in range(4)]
    dims = (48, 48, 48)
    # the synthetic weights are centred on latent row 0 only: take every code's own median value as its level
    levels = [float(np.median(R.decode_shape(c, dec, dims, use_net=1, padding=0.1, reshape=False, sigmoid=False))) for c in codes]
    serial = [R.create_mesh(c, dec, 1, sigmoid=False, level=lv, gradient_direction="ascent", dims=dims) for c, lv in zip(codes, levels)]
    pts = synth.make_points(50_000, 3).numpy().astype(np.float64)
    serial_vals = [R.decode_samples(c, dec, pts, use_net=1, sigmoid=False) for c in codes]
    results, values, errs = [None] * 4, [None] * 4, []

    def work(i):
        try:
            torch.cuda.set_device(0)
            for _ in range(5):
                results[i] = R.create_mesh(codes[i], dec, 1, sigmoid=False, level=levels[i], gradient_direction="ascent", dims=dims)
                values[i] = R.decode_samples(codes[i], dec, pts, use_net=1, sigmoid=False)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for i in range(4):
        assert np.array_equal(results[i].vertices, serial[i].vertices) and np.array_equal(results[i].faces, serial[i].faces)
        assert np.array_equal(values[i], serial_vals[i])
