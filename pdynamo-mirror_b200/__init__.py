"""pdynamo-mirror_b200: B200-native NBModelABFS hot path behind pDynamo's NB model plugin surface.
Import as `pdynamo_mirror_b200` through the loader shim at the repository root."""
from . import md, mmterms, parallel, workloads
from ._lib import CLibraryError, LIB_PATH
from .mmterms import FourierDihedralContainer, HarmonicAngleContainer, HarmonicBondContainer, HarmonicImproperContainer, MMTermsB200
from .nbmodel import NBModel, NBModelABFS, NBModelABFSState, PairListGenerator, PairwiseInteractionABFS, QCMMInteractionState
from .system import (Configuration, EnergyModel, LJParameterContainer, MMAtomContainer, SelfPairList, Symmetry,
                     SymmetryParameterGradients, SymmetryParameters, System, Transformation3Container)
