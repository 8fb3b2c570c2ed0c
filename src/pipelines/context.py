"""`src.pipelines.context` of the reference -> mikudance_b200.context."""
from mikudance_b200.context import get_context_scheduler, get_total_steps, ordered_halving, uniform  # noqa: F401
