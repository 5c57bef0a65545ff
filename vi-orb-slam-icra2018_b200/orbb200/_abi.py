"""ctypes signatures of include/orbb200.h (kept in the same order as the header)."""
import ctypes as C

vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

SIGNATURES = {
    "orb_last_error": (C.c_char_p, []),
    "orb_version": (i32, []),
    "orb_device_count": (i32, []),
    # extractor
    "orbx_create": (i32, [i32, f32, i32, i32, i32, i32, i32, i32, i32, vp]),
    "orbx_destroy": (i32, [vp]),
    "orbx_keypoint_capacity": (i32, [vp, vp]),
    "orbx_extract": (i32, [vp, vp, i32, i32, i32, vp, vp, i32, vp]),
    "orbx_extract_batch": (i32, [vp, vp, i32, i32, i32, i32, sz, vp, vp, i32, vp]),
    "orbx_extract_batch_device": (i32, [vp, vp, i32, i32, i32, i32, sz, vp, vp, i32, vp, vp]),
    "orbx_synchronize": (i32, [vp]),
    "orbx_get_levels": (i32, [vp, vp]),
    "orbx_get_scale_tables": (i32, [vp, vp, vp, vp, vp, vp, vp]),
    "orbx_get_level": (i32, [vp, i32, i32, vp, vp, vp]),
    "orbx_stage_times": (i32, [vp, vp]),
    "orbx_compute_stereo_matches": (i32, [vp, i32, vp, i32, vp, vp, i32, vp, vp, i32, f32, f32, vp, vp, vp]),
    "orbx_compute_stereo_matches_device": (i32, [vp, i32, vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, f32, f32, vp, vp,
                                                 vp, vp, vp]),
    "orbx_debug_candidates": (i32, [vp, i32, i32, vp, i32, vp]),
    "orbx_debug_blurred": (i32, [vp, i32, i32, vp]),
    "orbx_last_device_outputs": (i32, [vp, i32, vp, vp, vp, vp, vp]),
    "orbx_last_launch_count": (i32, [vp, vp]),
    "orbx_set_profiling": (i32, [vp, i32]),
    "orbx_kernel_times": (i32, [vp, vp, vp]),
    # matcher
    "orbm_create": (i32, [i32, vp]),
    "orbm_destroy": (i32, [vp]),
    "orbm_synchronize": (i32, [vp]),
    "orbm_last_launch_count": (i32, [vp, vp]),
    "orbm_distance": (i32, [vp, vp, vp, i32, vp]),
    "orbm_frame_create": (i32, [vp, vp, vp, i32, f32, f32, f32, f32, vp]),
    "orbm_frame_destroy": (i32, [vp]),
    "orbm_undistort_points": (i32, [vp, vp, vp, i32, vp]),
    "orbm_image_bounds": (i32, [vp, vp, i32, i32, vp]),
    "orbm_frame_create_device": (i32, [vp, vp, vp, vp, i32, vp, f32, f32, f32, f32, vp, vp]),
    "orbm_frame_size": (i32, [vp, vp]),
    "orbm_frame_download": (i32, [vp, vp, vp]),
    "orbm_frame_grid": (i32, [vp, vp, vp]),
    "orbm_features_in_area": (i32, [vp, vp, i32, i32, i32, vp, i32, vp]),
    "orbm_search_for_initialization": (i32, [vp, vp, vp, vp, vp, i32, f32, i32, vp]),
    "orbm_search_by_projection": (i32, [vp, vp, vp, i32, vp, f32, vp, vp, i32, f32, i32, vp, vp, i32, vp]),
    "orbm_search_by_projection_ex": (i32, [vp, vp, vp, i32, vp, f32, vp, vp, i32, f32, i32, i32, vp, vp, i32, vp]),
    "orbm_search_by_projection_world": (i32, [vp, vp, vp, i32, vp, f32, vp, vp, vp, i32, f32, i32, i32, vp, vp, i32, vp]),
    "orbm_project_points": (i32, [vp, vp, vp, i32, vp, vp, vp]),
    "orbm_search_by_projection_batch": (i32, [vp, vp, i32, vp, i32, f32, f32, i32, i32, i32, vp]),
    "orbm_search_for_initialization_batch": (i32, [vp, vp, i32, i32, f32, i32, vp]),
    "orbm_search_by_projection_points": (i32, [vp, vp, vp, i32, vp, vp, vp, i32, f32, f32, vp, vp, vp]),
    "orbm_search_for_triangulation": (i32, [vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, f32, f32,
                                            vp, vp, i32, i32, i32, vp, vp]),
    "orbm_search_by_bow": (i32, [vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, f32, i32, i32, vp, vp, vp]),
    "orbm_search_projected_best": (i32, [vp, vp, vp, vp, i32, i32, vp, vp, i32, vp, vp]),
    "orbm_bruteforce":(i32, [vp, vp, vp, i32, vp, vp, i32, i32, f32, i32, vp, vp, vp, vp, vp]),
    "orbm_bruteforce_device": (i32, [vp, vp, vp, i32, vp, vp, i32, i32, f32, i32, vp, vp, vp, vp, vp, vp]),
    "orbm_allpairs_device": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, i32, vp, vp]),
    "orbm_comm_unique_id": (i32, [vp]),
    "orbm_comm_create": (i32, [vp, i32, i32, i32, vp]),
    "orbm_comm_destroy": (i32, [vp]),
    "orbm_comm_info": (i32, [vp, vp, vp, vp]),
    "orbm_allpairs_sharded": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, f32, i32, vp, vp]),
    "orbm_comm_last_gather": (i32, [vp, vp, vp, vp]),
    "orbm_distinctive_descriptors": (i32, [vp, vp, vp, i32, vp, vp]),
    "orbm_vocabulary_create": (i32, [vp, i32, i32, vp, vp, vp, vp, vp, vp]),
    "orbm_vocabulary_destroy": (i32, [vp]),
    "orbm_bow_transform": (i32, [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "orbm_bow_transform_device": (i32, [vp, vp, vp, i32, i32, vp, vp, vp, vp]),
    "orbm_popc_peak": (i32, [vp, vp]),
}


def declare(lib):
    missing = []
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    return missing


class ProjectionJob(C.Structure):      # orbm_projection_job
    _fields_ = [("cur", vp), ("queries", vp), ("query_desc", vp), ("nq", i32), ("u_right", vp), ("occupied", vp),
                ("cur_match", vp), ("nmatches", i32)]


class InitJob(C.Structure):            # orbm_init_job
    _fields_ = [("f1", vp), ("f2", vp), ("prev_xy", vp), ("matches12", vp), ("nmatches", i32)]
