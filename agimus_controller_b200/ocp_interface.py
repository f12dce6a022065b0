"""Host-side mirror of the reference's OCP interface types (same names and field meaning).

The reference modules cannot be imported without Pinocchio (``agimus_controller/trajectory.py:5``), so the
few types the solve path touches are restated here, duck-compatible with the originals:

* ``OCPBase``                  ``agimus_controller/agimus_controller/ocp_base.py:11-107``
* ``OCPResults/OCPDebugData``  ``agimus_controller/agimus_controller/mpc_data.py:8-31``
* ``TrajectoryPoint``, ``TrajectoryPointWeights``, ``WeightedTrajectoryPoint``
                               ``agimus_controller/agimus_controller/trajectory.py:10-181``
* ``DTFactorsNSeq``, ``OCPParamsBaseCroco``
                               ``agimus_controller/agimus_controller/ocp_param_base.py:6-85``
* ``SE3``                      the two attributes of ``pinocchio.SE3`` the path reads (rotation, translation)

Objects of the reference's own classes are accepted wherever these are (attribute access only).
"""
from __future__ import annotations

import abc
import dataclasses
import typing as T

import numpy as np


@dataclasses.dataclass
class SE3:
    rotation: np.ndarray = dataclasses.field(default_factory=lambda: np.eye(3))
    translation: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(3))


@dataclasses.dataclass
class TrajectoryPoint:
    id: T.Optional[int] = None
    time_ns: T.Optional[int] = None
    robot_configuration: T.Optional[np.ndarray] = None
    robot_velocity: T.Optional[np.ndarray] = None
    robot_acceleration: T.Optional[np.ndarray] = None
    robot_effort: T.Optional[np.ndarray] = None
    forces: T.Optional[dict] = None
    end_effector_poses: T.Optional[dict] = None
    end_effector_velocities: T.Optional[dict] = None

    @property
    def robot_state(self) -> np.ndarray:
        return np.concatenate((self.robot_configuration, self.robot_velocity))


@dataclasses.dataclass
class TrajectoryPointWeights:
    w_robot_configuration: T.Optional[np.ndarray] = None
    w_robot_velocity: T.Optional[np.ndarray] = None
    w_robot_acceleration: T.Optional[np.ndarray] = None
    w_robot_effort: T.Optional[np.ndarray] = None
    w_forces: T.Optional[dict] = None
    w_end_effector_poses: T.Optional[dict] = None
    w_end_effector_velocities: T.Optional[dict] = None
    w_collision_avoidance: T.Optional[float] = None

    @property
    def w_robot_state(self) -> np.ndarray:
        return np.concatenate((self.w_robot_configuration, self.w_robot_velocity))


@dataclasses.dataclass
class WeightedTrajectoryPoint:
    point: TrajectoryPoint
    weights: TrajectoryPointWeights


@dataclasses.dataclass
class OCPResults:
    states: list = dataclasses.field(default_factory=list)
    ricatti_gains: list = dataclasses.field(default_factory=list)
    feed_forward_terms: list = dataclasses.field(default_factory=list)


@dataclasses.dataclass
class OCPDebugData:
    result: OCPResults = dataclasses.field(default_factory=OCPResults)
    references: list = dataclasses.field(default_factory=list)
    residuals: list = dataclasses.field(default_factory=list)
    kkt_norm: float = 0.0
    nb_iter: int = 0
    nb_qp_iter: int = 0
    problem_solved: bool = False


@dataclasses.dataclass
class DTFactorsNSeq:
    factors: list
    n_steps: list


@dataclasses.dataclass
class OCPParamsBaseCroco:
    dt: float
    solver_iters: int
    dt_factor_n_seq: DTFactorsNSeq
    horizon_size: int
    qp_iters: int = 200
    termination_tolerance: float = 1e-3
    max_solve_time: T.Optional[float] = None
    eps_abs: float = 1e-6
    eps_rel: float = 0.0
    callbacks: bool = False
    use_debug_data: bool = True
    n_threads: int = 1
    use_filter_line_search = False

    def __post_init__(self):
        seq = self.dt_factor_n_seq
        self._n_controls = int(sum(seq.n_steps))
        steps: list = []
        for factor, n in zip(seq.factors, seq.n_steps):
            steps += [self.dt * factor] * n
        self.timesteps = tuple(steps)
        self.total_time = sum(steps)
        assert self.horizon_size == self._n_controls, (
            f"The horizon size {self.horizon_size} must be equal to the sum of the time steps {self._n_controls}.")

    @property
    def n_controls(self) -> int:
        return self._n_controls


class OCPBase(abc.ABC):
    """The interface ``MPC.run`` drives (``agimus_controller/agimus_controller/mpc.py:32-66``)."""

    @abc.abstractmethod
    def set_reference_weighted_trajectory(self, reference_weighted_trajectory: list) -> None: ...

    @property
    @abc.abstractmethod
    def n_controls(self) -> int: ...

    @property
    def horizon_size(self) -> int:
        return self.n_controls

    @property
    @abc.abstractmethod
    def dt(self) -> float: ...

    @abc.abstractmethod
    def solve(self, x0, x_warmstart, u_warmstart, use_iteration_limits_and_timeout: bool = True) -> None: ...

    @abc.abstractmethod
    def integrate(self, state, control): ...

    @property
    @abc.abstractmethod
    def ocp_results(self) -> OCPResults: ...

    @property
    @abc.abstractmethod
    def debug_data(self) -> OCPDebugData: ...
