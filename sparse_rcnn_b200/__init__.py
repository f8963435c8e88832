"""sparse_rcnn_b200: B200-native (sm_100a) sparse submanifold-convolution hot path behind the
`sparseconvnet` API surface that LeonhardFeiner/sparse_rcnn consumes.  See DESIGN.md."""
import sys

__version__ = "0.1.0"


def install_as_sparseconvnet():
    """Alias `sparse_rcnn_b200.scn` as `sparseconvnet` so `import sparseconvnet as scn` in the
    reference's ndsis/modules binds to the B200 kernels (the drop-in boundary)."""
    from . import scn
    sys.modules["sparseconvnet"] = scn
    return scn
