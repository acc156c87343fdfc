"""diffusers.callbacks: names imported by the reference pipelines for type hints only."""
from .models._placeholder import placeholder

PipelineCallback = placeholder("PipelineCallback")
MultiPipelineCallbacks = placeholder("MultiPipelineCallbacks")
