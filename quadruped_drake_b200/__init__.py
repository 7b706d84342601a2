"""B200-native batched whole-body QP controller (drop-in for the per-step path of
vincekurtz/quadruped_drake's controllers: BasicController, IDController, CLFController, MPTCController, PCController)."""
from .model import RobotModel, load_robot, flatten  # noqa: F401

__all__ = ["RobotModel", "load_robot", "flatten"]
