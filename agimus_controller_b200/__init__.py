"""B200-native batched OCP solve path behind agimus_controller's ``OCPBase`` interface."""
from .robot_model import RobotTable, panda_table, PANDA_Q_NOMINAL  # noqa: F401
