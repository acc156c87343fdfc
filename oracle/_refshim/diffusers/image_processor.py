"""diffusers.image_processor: imported by the reference pipelines; the T2V pipeline never instantiates it."""
from .models._placeholder import placeholder

VaeImageProcessor = placeholder("VaeImageProcessor")
