"""`src.pipelines.pipeline_stage2_vdo` of the reference -> mikudance_b200.pipelines."""
from mikudance_b200.pipelines import Pose2VideoPipeline, Pose2VideoPipelineOutput  # noqa: F401
