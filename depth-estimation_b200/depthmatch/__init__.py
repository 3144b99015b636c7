"""depthmatch -- host-side mirror of the reference's Lua surface for the dense
matching hot path, on top of libdepthmatch.so (hand-written sm_100a CUDA).

The reference is Torch7/Lua and no Lua interpreter exists in this image, so the
tested host layer is this Python package; the LuaJIT-FFI shims with the same
names live in ../lua/ (see INTEGRATION.md).  Function and class names, argument
meaning and 1-based index conventions follow the reference:

    nn.SpatialMatching / nn.SpatialRadialMatching      (out-of-tree nnx)
    nn.CascadingAddTable, nn.OutputExtractor            CascadingAddTable.lua, OutputExtractor.lua
    extractoutput.extractOutput[Marginalized]           extract_output.cpp
    prepareInput / getModel / processOutput / x2yx ...  opticalflow_model.lua
    getModelMultiscale / yx2xMulti / x2yxMulti          opticalflow_model_multiscale.lua
    getC2PMask / getP2CMask / cartesian2polar / getRMax radial/*.lua

There is no CPU fallback: everything below calls the CUDA library.
"""
from ._lib import DepthMatchError, load  # noqa: F401
from .api import *  # noqa: F401,F403
from . import torch7io  # noqa: F401
from .model_io import (loadCalibration, loadModel, loadTesterNetwork, loadWeightsFrom,  # noqa: F401
                       modelDirectory, saveModel, saveNetwork)
from .depth_estimation_api import DepthEstimationAPI, removeEgoMotion, warpHomography  # noqa: F401
from .radial_api import RadialTester, polarGeometry  # noqa: F401
from .stream import FeatureStream  # noqa: F401
