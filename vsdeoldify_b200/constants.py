"""Default parameter lists of the hot path — the values (not the code) of vsdeoldify/vsslib/constants.py:13-57,
which HAVC_colorizer / HAVC_merge take positionally."""
DEF_LEVEL_NONE = 0
DEF_LEVEL_INFO = 1
DEF_LEVEL_DEBUG = 2
DEF_CMC_p = [0.15, True, 20, 24]            # ConstrainedChromaMerge: chroma threshold, red-fix, ...
DEF_LMM_p = [0.15, 0.65, 1.0]               # LumaMaskedMerge: luma limit, white limit, alpha
DEF_ALM_p = [0.8, 1.0, 0.15]                # AdaptiveLumaMerge: luma threshold, alpha, min weight
DEF_CRT_p = [0.8, 30, 2, False, 0, 0]       # ChromaRetentionMerge
DEF_TWEAK_p = [0.0, 1.0, 2.5, True, 0.3, 0.6, 1.5, 0.5]
DEF_THT_WHITE = 0.70
DEF_THT_BLACK = 0.10
DEF_STANDARD_DARK = 0.22
DEF_STANDARD_BRIGHT = 0.78
DEF_STABLE_WEIGHT = 0.50                    # video_weight of the 'stable' blend (vsmodels.py:204-206)
DEF_ARTISTIC_WEIGHT = 0.50                  # video_weight of the 'artistic' blend (vsmodels.py:207-209)
