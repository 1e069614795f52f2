from wtracker_b200.sim.config import ExperimentConfig, TimingConfig
from wtracker_b200.sim.motor_controllers import MotorController, SineMotorController, StepMotorController
from wtracker_b200.sim.simulator import SimController, Simulator
from wtracker_b200.sim.view_controller import ViewController

__all__ = ["ExperimentConfig", "TimingConfig", "MotorController", "SineMotorController", "StepMotorController",
           "SimController", "Simulator", "ViewController"]
