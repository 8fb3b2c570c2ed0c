"""`src.pipelines.pipeline_mikudance` of the reference -> mikudance_b200.pipelines."""
from mikudance_b200.pipelines import MikuDanceVideoPipeline, MikuDanceVideoPipelineOutput  # noqa: F401
