"""`sparse_rcnn_b200.scn` -- drop-in for the `sparseconvnet` namespace that
LeonhardFeiner/sparse_rcnn's ndsis/modules import as `scn` (module_factory.py:5, model.py:6,
custom_operations.py:4, roi_select_sparse.py:3).  `sparse_rcnn_b200.install_as_sparseconvnet()`
aliases it in sys.modules so the reference modules run on the B200 kernels unchanged.
"""
import types

from .functions import (InputLayerFunction, OutputLayerFunction, get_precision, set_precision)
from .layers import (AddTable, AveragePooling, BatchNormalization, BatchNormLeakyReLU, BatchNormReLU, ConcatTable,
                     Convolution, Deconvolution, Identity, InputLayer, JoinTable, MaxPooling, NetworkInNetwork,
                     OutputLayer, ReLU, Sequential, SparseConvNetTensor, SparseToDense, SubmanifoldConvolution,
                     UnPooling, ValidConvolution)
from .metadata import GeometryPrefetcher, Metadata, get_row_order, set_row_order

ioLayers = types.SimpleNamespace(
    InputLayerFunction=InputLayerFunction, OutputLayerFunction=OutputLayerFunction,
    InputLayer=InputLayer, OutputLayer=OutputLayer)

BACKEND = "b200-cuda"
