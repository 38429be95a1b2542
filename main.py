"""Entry point: the statements of the reference's main.py (main.py:4-41) with its own imports -- `sydr` is the import
surface of this repository (sydr/__init__.py: every `sydr.X` is `sydr_b200.X`; the terminal GUI, the HTML report and
the logging set-up of the reference are outside the hot path and are no-op stand-ins).  Two additions: the
configuration file may be given as first argument, and `--fast` uses the whole-file streaming path.

    python main.py [config/receiver.ini] [--fast]
"""
import configparser
import sys

from sydr.enlightengui import EnlightenGUI
from sydr.receiver.receiver_gps_l1ca import ReceiverGPSL1CA
from sydr.io.visualisation import Visualisation

import sydr.logger as logger


def main():
    # Configuration
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    receiverConfigFile = args[0] if args else './config/receiver.ini'
    receiverConfig = configparser.ConfigParser()
    if not receiverConfig.read(receiverConfigFile):
        raise SystemExit(f"cannot read {receiverConfigFile}")

    gui = EnlightenGUI()
    gui.updateMainStatus(stage='Initialize', status='RUNNING')
    logger.configureLogger(name=__name__, filepath='./config/logging.ini')

    receiver = ReceiverGPSL1CA(receiverConfig, overwrite=True, gui=gui)
    if "--fast" in sys.argv:
        receiver.run_fast()
    else:
        receiver.run()
    receiver.close()

    gui.updateMainStatus(stage='Create report', status='RUNNING')
    Visualisation(receiverConfig).run()
    gui.updateMainStatus(stage='PROCESSING COMPLETED', status='DONE')


if __name__ == "__main__":
    main()
