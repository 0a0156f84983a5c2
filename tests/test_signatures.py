"""CPU: the drop-in modules export the reference's names with the reference's positional signatures and defaults.

The reference source is parsed (never imported) when it is present in this container; the expected signatures
are also pinned below, so the test still means something on a machine without /root/reference."""
import ast
import inspect
import os

import pytest

REF = "/root/reference/2cam"

# name -> (positional parameter names, defaults of the trailing ones), from 2cam/vision_pipeline_utils.py:12-381
VPU = {
    "erode_mask_gpu": (["mask", "kernel_size"], [3]),
    "convert_mask_to_3d_points": (["mask_indices", "depth_map", "cx", "cy", "fx", "fy"], []),
    "downsample_point_cloud_gpu": (["point_cloud", "voxel_size"], []),
    "filter_outliers_sor": (["point_cloud", "nb_neighbors", "std_ratio"], [20, 1.5]),
    "point_clouds_equal": (["pc1", "pc2"], []),
    "visualize_point_cloud": (["point_cloud", "title"], ["Point Cloud"]),
    "calculate_centroid": (["point_cloud"], []),
    "crop_point_cloud_gpu": (["point_cloud", "x_bounds", "y_bounds", "z_bounds"], []),
    "fuse_point_clouds_centroid": (["point_clouds_camera1", "point_clouds_camera2", "distance_threshold"], [0.1]),
    "subtract_point_clouds_gpu": (["workspace_pc", "objects_pc", "distance_threshold"], [0.005]),
    "retrieve_frames": (["zed1", "zed2", "image1", "image2", "timings"], []),
    "retrieve_depth_maps": (["zed1", "zed2", "depth1", "depth2", "timings"], []),
    "process_point_clouds": (["zed1", "zed2", "point_cloud1_ws", "point_cloud2_ws", "resolution", "rotation1_torch", "origin1_torch",
                              "rotation2_torch", "origin2_torch", "device", "timings"], []),
    "perform_yolo_inference": (["model", "frame1", "frame2", "device", "timings"], []),
    "process_masks": (["masks", "class_ids", "zed_depth_np", "cx", "cy", "fx", "fy", "rotation_robot_cam_torch", "origin_cam_torch",
                       "device", "point_clouds_camera"], []),
    "fuse_point_clouds": (["point_clouds_camera1", "point_clouds_camera2", "distance_threshold", "timings"], []),
    "subtract_point_clouds": (["fused_point_cloud_ws", "fused_pc_objects_concatenated", "distance_threshold", "timings"], []),
}
# from 2cam/2cams_mask_cpu.py:47-309
MASK_CPU = {
    "erode_mask_gpu": (["mask", "kernel_size"], [12]),
    "erode_mask": (["mask", "iterations"], [1]),
    "downsample_point_cloud": (["point_cloud", "voxel_size"], [0.01]),
    "convert_mask_to_3d_points_gpu": (["mask_indices", "depth_map", "cx", "cy", "fx", "fy"], []),
    "convert_mask_to_3d_points_cpu": (["mask_indices", "depth_map", "cx", "cy", "fx", "fy", "device"], ["cuda"]),
    "downsample_point_cloud_gpu": (["point_cloud", "voxel_size"], []),
    "filter_outliers_sor": (["point_cloud", "nb_neighbors", "std_ratio"], [20, 1.5]),
    "point_clouds_equal": (["pc1", "pc2"], []),
    "visualize_point_cloud": (["point_cloud", "title"], ["Point Cloud"]),
    "calculate_centroid": (["point_cloud"], []),
    "calculate_centroid_gpu": (["point_cloud"], []),
    "crop_point_cloud_gpu": (["point_cloud", "x_bounds", "y_bounds", "z_bounds"], []),
    "fuse_point_clouds_centroid": (["point_clouds_camera1", "point_clouds_camera2", "distance_threshold"], [0.1]),
    "subtract_point_clouds": (["workspace_pc", "objects_pc", "distance_threshold"], [0.01]),
    "subtract_point_clouds_gpu": (["workspace_pc", "objects_pc", "distance_threshold"], [0.005]),
    "voxel_grid_subtract": (["workspace_pc", "objects_pc", "voxel_size"], [0.02]),
}


def ours(mod, name):
    sig = inspect.signature(getattr(mod, name))
    params = list(sig.parameters.values())
    return [p.name for p in params], [p.default for p in params if p.default is not inspect.Parameter.empty]


def reference_signatures(path):
    tree = ast.parse(open(path).read())
    out = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            out[node.name] = ([a.arg for a in node.args.args], [ast.literal_eval(d) for d in node.args.defaults])
    return out


@pytest.mark.parametrize("modname,table", [("vision_pipeline_utils", VPU), ("compat_mask_cpu", MASK_CPU)])
def test_names_positions_and_defaults(modname, table):
    mod = __import__(modname)
    for name, (args, defaults) in table.items():
        assert hasattr(mod, name), f"{modname} lacks {name}"
        got_args, got_defaults = ours(mod, name)
        assert got_args[:len(args)] == args, f"{modname}.{name}: positional parameters {got_args} != {args}"
        assert got_defaults[:len(defaults)] == defaults, f"{modname}.{name}: defaults {got_defaults} != {defaults}"
        # anything we add must be optional, so that the reference's call sites keep working
        sig = inspect.signature(getattr(mod, name))
        for p in list(sig.parameters.values())[len(args):]:
            assert p.default is not inspect.Parameter.empty


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present on this machine")
@pytest.mark.parametrize("fname,table", [("vision_pipeline_utils.py", VPU), ("2cams_mask_cpu.py", MASK_CPU)])
def test_pinned_tables_match_the_reference_source(fname, table):
    ref = reference_signatures(os.path.join(REF, fname))
    for name, expected in table.items():
        assert name in ref, f"{fname} has no {name}"
        assert ref[name] == (expected[0], expected[1]), f"{fname}:{name} is {ref[name]}, table says {expected}"
    if fname == "vision_pipeline_utils.py":   # the star import must be complete
        missing = sorted(set(ref) - set(table) - {"process_iteration_end"})
        assert not missing, f"reference functions without a pinned signature: {missing}"


def test_star_import_is_complete():
    import vision_pipeline_utils as vpu
    for name in list(VPU) + ["process_iteration_end", "np", "cv2", "time", "torch", "F", "csv", "sl", "o3d", "YOLO"]:
        assert hasattr(vpu, name), name
