from .._placeholder import placeholder

CogVideoXCausalConv3d = placeholder("CogVideoXCausalConv3d")
CogVideoXDownBlock3D = placeholder("CogVideoXDownBlock3D")
CogVideoXMidBlock3D = placeholder("CogVideoXMidBlock3D")
CogVideoXSafeConv3d = placeholder("CogVideoXSafeConv3d")
CogVideoXSpatialNorm3D = placeholder("CogVideoXSpatialNorm3D")
CogVideoXUpBlock3D = placeholder("CogVideoXUpBlock3D")
