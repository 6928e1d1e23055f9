"""driving-dirty_b200: the six-camera scene pipeline of annikabrundyn/driving-dirty, rebuilt for
B200 (sm_100a).  Same class names, constructor/forward signatures and state_dict keys as the
reference's modules for this path; the arithmetic runs in libdd_b200.so (hand-written CUDA),
reached through the C ABI in include/dd_b200.h.

    from driving_dirty_b200.autoencoder.autoencoder import BasicAE
    from driving_dirty_b200.roadmap_model.roadmap_bce_v2 import RoadMapBCE
    from driving_dirty_b200.utils.helper import compute_ts_road_map, collate_fn
    from driving_dirty_b200.model_loader import ModelLoader
"""
__version__ = "0.1.0"
