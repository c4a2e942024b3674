"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.

CPU restatement of the per-frame filter -> segment hot path of
david-zwicker/video-analysis.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
package, and only as the checker / the timed CPU baseline.  The product package
``video_analysis_b200`` never imports it and has no CPU fallback.

The reference itself cannot be imported (Python 2 only, un-vendored ``utils``
package, ``shapely`` missing -- SURVEY.md section 8c), so every function here
issues *the same library call at the same call site* (cv2 / NumPy / SciPy, all
present in the image) and cites the reference ``file:line`` it follows.

PARITY STATUS
  * monochrome, crop, blur, resize, label, erode/dilate: pinned by calling the
    identical library function with identical arguments as the cited line.
    The reference has no tests or golden vectors, so nothing else can pin them.
  * background EMA, threshold, open/close, apply-mask: these ops do not exist
    in the reference snapshot ("parity unpinned" by the reference).  The
    definitions adopted here are the ones SURVEY.md section 8c states, built
    from the nearest reference idioms (cited per function).
"""

from . import ops, synth  # noqa: F401
