from ._placeholder import placeholder

AutoencoderKL = placeholder("AutoencoderKL")          # imported for type hints by the reference pipelines
HunyuanDiT2DModel = placeholder("HunyuanDiT2DModel")
