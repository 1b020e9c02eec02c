"""Public names (the reference module exports nothing and is used as `ProximalAlgorithms.X`; same here: `proxb200.X`)."""
from ._lib import LIB_PATH, ProxB200Error  # noqa: F401
from .algorithms import (  # noqa: F401
    FastForwardBackward,
    FastForwardBackwardIteration,
    FastForwardBackwardState,
    FastProximalGradient,
    FastProximalGradientIteration,
    ForwardBackward,
    ForwardBackwardIteration,
    ForwardBackwardState,
    IterativeAlgorithm,
    ProximalGradient,
    ProximalGradientIteration,
    default_display,
    default_solution,
    default_stopping_criterion,
)
from .accel import LBFGS, LBFGSOperator, NoAcceleration  # noqa: F401
from .douglas_rachford import DouglasRachford, DouglasRachfordIteration, DouglasRachfordState  # noqa: F401
from .functions import (  # noqa: F401
    BlockDiagLeastSquares,
    FiniteDifference2D,
    IndBallL2,
    IndBox,
    LeastSquares,
    LinearFunction,
    MatrixOp,
    NormL1,
    NormL21,
    SqrNormL2,
    SquaredDistance,
    Zero,
)
from .primal_dual import AFBA, AFBAIteration, ChambollePock, ChambollePockIteration, VuCondat, VuCondatIteration  # noqa: F401
from .tv import IndConsensus, TVSplit  # noqa: F401
from .panoc import PANOC, PANOCIteration, PANOCState  # noqa: F401
from . import iteration_tools as IterationTools  # noqa: F401
from .jld2 import load_lasso_fixture, read_jld2  # noqa: F401
from .host import Context, DeviceExchangeComm, LocalComm, Scalars, TorchDistComm, dense_shard_bounds, shard_bounds  # noqa: F401
from .nesterov import (  # noqa: F401
    AdaptiveNesterovSequence,
    ConstantNesterovSequence,
    FixedNesterovSequence,
    SimpleNesterovSequence,
)

__all__ = [n for n in dir() if not n.startswith("_")]
