"""Drop-in surface audit: every attribute the reference's own callers read on the model / renderer objects
exists on the B200 mirrors.  Reads the reference sources, so it runs in the build container only (skipped elsewhere)."""
import os
import re

import pytest

import contrastive_lift_b200 as cl
from contrastive_lift_b200 import synthetic as syn

REF = os.environ.get("CLIFT_REFERENCE_ROOT", "/root/reference")
CALLERS = ["trainer/train_panopli_tensorf.py", "inference/render_panopli.py", "inference/extract_train_centroids.py",
           "inference/find_bandwidth.py", "inference/render_panopli_original.py"]
# attributes that only exist behind switches this repo refuses loudly, or in commented-out code
OUT_OF_SCOPE = {"model": {"proj_layer", "tv_loss_distilled_features"},           # use_proj, distilled-feature grids
                "renderer": {"export_instance_clusters"}}                          # visualisation, commented out at trainer:411

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "trainer")), reason="reference tree not present")


def used_attributes():
    pat = re.compile(r"(?<![\w.])(?:self\.)?(model|renderer)\.([A-Za-z_]\w*)")
    found = {"model": set(), "renderer": set()}
    for rel in CALLERS:
        for line in open(os.path.join(REF, rel)):
            code = line.split("#", 1)[0]
            if code.lstrip().startswith(("from ", "import ")):
                continue
            for obj, attr in pat.findall(code):
                found[obj].add(attr)
    return found


def test_every_attribute_the_reference_callers_use_exists():
    grid = [8, 8, 8]
    model = cl.TensorVMSplit(grid, num_semantics_comps=(32, 32, 32), num_instance_comps=(32, 32, 32), num_semantic_classes=4,
                             dim_feature_instance=6, use_semantic_mlp=True, use_instance_mlp=True, slow_fast_mode=True)
    rend = cl.TensoRFRenderer(syn.default_aabb(), grid, semantic_weight_mode="softmax")
    used = used_attributes()
    assert {"upsample_volume_grid", "get_optimizable_parameters", "total_tv_loss", "slow_fast_mode"} <= used["model"]
    assert {"forward_instance_feature", "update_step_ratio", "update_bbox_aabb_and_shrink", "grid_dim"} <= used["renderer"]
    missing = {(o, a) for o, obj in (("model", model), ("renderer", rend)) for a in used[o] - OUT_OF_SCOPE[o] if not hasattr(obj, a)}
    assert not missing, f"reference callers use attributes the mirrors lack: {sorted(missing)}"
    # the nested ones the EMA reads (trainer:259)
    assert hasattr(model.render_instance_mlp, "mlp") and hasattr(model.render_instance_mlp, "slow_mlp")
