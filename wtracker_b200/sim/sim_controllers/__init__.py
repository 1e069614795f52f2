from wtracker_b200.sim.sim_controllers.csv_controller import CsvController
from wtracker_b200.sim.sim_controllers.logging_controller import LogConfig, LoggingController
from wtracker_b200.sim.sim_controllers.mlp_controllers import MLPController
from wtracker_b200.sim.sim_controllers.yolo_controller import YoloConfig, YoloController

__all__ = ["CsvController", "LogConfig", "LoggingController", "MLPController", "YoloConfig", "YoloController"]
