"""GPU ingest stage: convert -> rectify -> back-project -> body-frame transform (-> gather)."""

from thor_slam_b200.ingest import formats
from thor_slam_b200.ingest.context import IngestContext, StreamSpec

__all__ = ["IngestContext", "StreamSpec", "formats"]
