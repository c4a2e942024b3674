"""
Generates tests/golden/golden_ref.npz by running THE REFERENCE'S OWN CODE (the Python-3 conversion
of /root/reference that oracle/build_ref.py writes into oracle/_ref) on seeded synthetic frames in
the build container (numpy 2.3.5 / OpenCV 4.13.0 / SciPy 1.18.1).  These vectors pin the oracle
(tests/test_oracle_vs_reference.py) and, through it and directly, the CUDA path
(tests/test_reference_golden_gpu.py).  /root/reference does not exist on the GPU box; the fixtures do.

    python tests/golden/make_golden_ref.py

What is stored (prefix -> reference code that produced it):
  f_*    video/filters.py     FilterCrop / FilterMonochrome / FilterBlur / FilterResize / FilterRotate /
                              FilterNormalize / FilterReplicate / FilterDropFrames / FilterTimeDifference
                              chains over video/io/memory.py VideoMemory, materialised with
                              VideoBase.copy() (video/io/base.py:248-269) or by iteration
  a_*    video/analysis       get_largest_region, find_bounding_box (regions.py:113-174),
                              measure_mean, measure_mean_std, reduce_video (video.py:14-55),
                              regionprops (image.py:310-405)
  p_*    video/io/base.py     protocol observations (strings, shapes, slice contents, cursor positions)
"""

import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref, synth  # noqa: E402


def stack(video):
    return np.stack([np.array(f) for f in video])


def reference_cases(ref):
    """ runs the reference; returns {name: ndarray}.  Shared with tests/test_oracle_vs_reference.py,
    which re-runs it live when oracle/_ref is importable and compares with the committed file. """
    F, M = ref.filters, ref.memory
    out = {}
    col = synth.make_frames(0, 0, 6, 64, 48, 4)            # (6, 48, 64, 3)
    rag = synth.make_frames(3, 10, 5, 53, 37, 3)           # ragged size
    out['f_col'] = col
    out['f_rag'] = rag

    def vm(a):
        return M.VideoMemory(a.copy())

    # --- filter classes -----------------------------------------------------------------------
    out['f_mono_mean'] = stack(F.FilterMonochrome(vm(col)))
    out['f_mono_green'] = stack(F.FilterMonochrome(vm(col), 'green'))
    out['f_mono_r'] = stack(F.FilterMonochrome(vm(rag), 'R'))
    out['f_crop_rect'] = stack(F.FilterCrop(vm(col), rect=(5, 7, 40, 30)))
    out['f_crop_frac'] = stack(F.FilterCrop(vm(col), rect=(0.25, 0.5, 0.5, 0.25)))
    out['f_crop_neg'] = stack(F.FilterCrop(vm(col), rect=(-30, -20, 17, 11)))
    out['f_crop_region'] = stack(F.FilterCrop(vm(rag), region='lower right'))
    out['f_crop_chan'] = stack(F.FilterCrop(vm(col), rect=(3, 2, 33, 21), color_channel='red'))
    out['f_crop_align'] = stack(F.FilterCrop(vm(col), rect=(1, 1, 37, 29), size_alignment=4))
    nested = F.FilterCrop(F.FilterCrop(vm(col), rect=(4, 6, 50, 40)), rect=(3, 5, 20, 10))
    out['f_crop_nested'] = stack(nested)
    out['p_crop_nested_rect'] = np.array(nested.rect)
    for s in (0.5, 1, 2, 3, 5):
        out['f_blur_%g' % s] = stack(F.FilterBlur(F.FilterMonochrome(vm(col)), sigma=s))
    out['f_blur_default_rag'] = stack(F.FilterBlur(F.FilterMonochrome(vm(rag))))
    out['f_blur_color'] = stack(F.FilterBlur(vm(col), sigma=1.5))
    mono = F.FilterMonochrome(vm(col))
    out['f_resize_half'] = stack(F.FilterResize(F.FilterMonochrome(vm(col)), 0.5))
    out['f_resize_third_even'] = stack(F.FilterResize(F.FilterMonochrome(vm(col)), 1 / 3., even_dimensions=True))
    out['f_resize_size'] = stack(F.FilterResize(F.FilterMonochrome(vm(col)), (40, 27)))
    out['f_resize_up_auto'] = stack(F.FilterResize(F.FilterMonochrome(vm(rag)), 1.5))
    for interp in ('nearest', 'linear', 'area', 'cubic', 'lanczos'):
        out['f_resize_%s' % interp] = stack(F.FilterResize(F.FilterMonochrome(vm(col)), (45, 31), interpolation=interp))
    out['f_resize_color_half'] = stack(F.FilterResize(vm(col), 0.5))
    out['f_resize_same'] = stack(F.FilterResize(mono, 1))
    for a in (90, 180, 270, 450):
        out['f_rot_%d' % a] = stack(F.FilterRotate(F.FilterMonochrome(vm(rag)), a))
    out['f_norm'] = stack(F.FilterNormalize(F.FilterMonochrome(vm(col)), 70, 180))
    out['f_norm_auto'] = stack(F.FilterNormalize(F.FilterMonochrome(vm(col))))
    out['f_norm_float'] = stack(F.FilterNormalize(F.FilterMonochrome(vm(col)), 60, 200, dtype=np.float32))
    # the reference's FilterReplicate.get_next_frame (filters.py:419-430) rewinds its source for ever and
    # never raises StopIteration: take frame_count frames, as VideoBase.copy()'s preallocated array would
    rep = F.FilterReplicate(F.FilterMonochrome(vm(rag)), 3)
    out['f_replicate'] = np.stack([np.array(f) for f in itertools.islice(rep, rep.frame_count)])
    out['f_drop'] = stack(F.FilterDropFrames(F.FilterMonochrome(vm(col)), 2))
    td = F.FilterTimeDifference(F.FilterMonochrome(vm(col)))
    out['f_timediff_get'] = np.stack([td.get_frame(i) for i in range(td.frame_count)])
    out['f_timediff_get_neg'] = td.get_frame(-1)
    # the headline front: crop -> mono -> blur -> resize, materialised by VideoBase.copy()
    chain = F.FilterResize(F.FilterBlur(F.FilterMonochrome(F.FilterCrop(vm(col), rect=(8, 4, 48, 40))), 2), 0.5)
    out['f_chain_copy'] = chain.copy(disp=False).data
    out['p_chain_str'] = np.array(str(chain))
    out['p_chain_shape'] = np.array(chain.shape)
    out['p_chain_format'] = np.array(json.dumps(chain.video_format, sort_keys=True))

    # --- analysis helpers ---------------------------------------------------------------------
    R, V, I = ref.regions, ref.video, ref.image
    rng = np.random.RandomState(7)
    masks = []
    yy, xx = np.mgrid[:60, :80]
    blobs = np.zeros((60, 80), bool)
    for cx, cy, r in ((12, 10, 6), (40, 30, 14), (70, 50, 8), (20, 45, 3)):
        blobs |= (xx - cx) ** 2 + (yy - cy) ** 2 <= r * r
    masks.append(blobs)
    masks.append(rng.rand(60, 80) > 0.55)
    u = np.zeros((60, 80), bool)
    u[5:50, 10:14] = u[5:50, 40:44] = u[46:50, 10:44] = True     # U shape: one region, two raster starts
    u[2:4, 20:30] = True
    masks.append(u)
    masks = np.stack(masks)
    out['a_masks'] = masks
    out['a_largest'] = np.stack([R.get_largest_region(m) for m in masks])
    out['a_largest_area'] = np.array([R.get_largest_region(m, ret_area=True)[1] for m in masks])
    out['a_largest_u8'] = R.get_largest_region(masks[0].astype(np.uint8) * 255)
    out['a_bbox'] = np.array([R.find_bounding_box(m) for m in out['a_largest']])
    props = [I.regionprops(m.astype(np.uint8)) for m in out['a_largest']]
    # (regionprops.eccentricity is not recorded: image.py:391 *calls* the inertia_tensor_eigvals property,
    # which image.py:397,402 read as an attribute -- it raises TypeError in the reference itself)
    out['a_props'] = np.array([[p.area, p.centroid[0], p.centroid[1], p.orientation,
                                p.inertia_tensor_eigvals[0], p.inertia_tensor_eigvals[1],
                                p.major_axis_length, p.minor_axis_length] for p in props])
    out['a_moments'] = np.array([[p.moments[k] for k in sorted(p.moments)] for p in props])
    out['a_moment_keys'] = np.array(sorted(props[0].moments))
    mv = F.FilterMonochrome(vm(col))
    out['a_mean'] = V.measure_mean(mv)
    m, s = V.measure_mean_std(F.FilterMonochrome(vm(col)))
    out['a_mean2'], out['a_std'] = m, s
    out['a_reduce_max'] = V.reduce_video(F.FilterMonochrome(vm(col)), np.maximum)
    out['a_reduce_sum_init'] = V.reduce_video(F.FilterMonochrome(vm(rag)), lambda f, acc: acc + f,
                                              np.zeros((37, 53), np.int64))

    # --- protocol observations ----------------------------------------------------------------
    v = vm(col)
    obs = []
    obs.append(str(v))
    obs.append(v.info())
    obs.append(repr((len(v), v.width, v.height, v.bounds, v.shape, v.is_color, v.fps)))
    flt = F.FilterMonochrome(F.FilterCrop(v, rect=(1, 2, 30, 20)))
    seen = []
    flt.register_listener(lambda f: seen.append(f.shape))
    obs.append(str(flt))
    fr = [flt.get_next_frame() for _ in range(2)]
    obs.append(repr((flt.get_frame_pos(), v.get_frame_pos(), seen)))
    flt.set_frame_pos(-2)
    obs.append(repr((flt.get_frame_pos(), v.get_frame_pos())))
    sl = flt[1:5:2]
    obs.append(str(sl))
    obs.append(repr((len(sl), sl.shape)))
    out['p_slice'] = stack(sl)
    out['p_slice_rev'] = stack(F.FilterMonochrome(vm(col))[4:0:-1])
    out['p_neg_index'] = F.FilterMonochrome(vm(col))[-1]
    for bad in (6, -7):
        try:
            vm(col).get_frame(bad)
            obs.append('no error')
        except Exception as e:                                     # noqa: BLE001
            obs.append(type(e).__name__)
    for args in (dict(rect=(0, 0, 64, 10)), dict(rect=(70, 0, 10, 10))):
        try:
            F.FilterCrop(vm(col), **args)
            obs.append('no error')
        except Exception as e:                                     # noqa: BLE001
            obs.append(type(e).__name__)
    for ctor in (lambda: F.FilterResize(vm(col), 0.5, interpolation='bogus'),
                 lambda: F.FilterRotate(vm(col), 45),
                 lambda: F.FilterMonochrome(vm(col), 1)):
        try:
            ctor()
            obs.append('no error')
        except Exception as e:                                     # noqa: BLE001
            obs.append(type(e).__name__)
    it = iter(F.FilterMonochrome(vm(col)))
    n = 0
    try:
        while True:
            next(it)
            n += 1
    except StopIteration:
        obs.append('StopIteration after %d' % n)
    out['p_obs'] = np.array(obs)
    return out


def main():
    ref = build_ref.import_ref()
    if ref is None:
        raise SystemExit('reference sources not found and oracle/_ref not built')
    out = reference_cases(ref)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden_ref.npz')
    np.savez_compressed(path, **out)
    print('%d arrays, %.1f kB' % (len(out), os.path.getsize(path) / 1e3))


if __name__ == '__main__':
    main()
