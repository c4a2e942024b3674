"""
video_analysis_b200 -- B200 (sm_100a) implementation of the per-frame
filter -> segment hot path of david-zwicker/video-analysis behind the reference's
own VideoBase / VideoFilterBase iteration API.
"""

__version__ = '0.1.0'
