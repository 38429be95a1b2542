"""Entry point with the shape of the reference's main.py: read ./config/receiver.ini (or the file given
as first argument), build the GPS L1 C/A receiver, run it, close it.  The GUI and the HTML report of
the reference are not part of this repository; `--fast` uses the whole-file streaming path.

    python main.py [config/receiver.ini] [--fast]
"""
import configparser
import sys

from sydr_b200.receiver.receiver_gps_l1ca import ReceiverGPSL1CA


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    receiverConfigFile = args[0] if args else './config/receiver.ini'
    receiverConfig = configparser.ConfigParser()
    if not receiverConfig.read(receiverConfigFile):
        raise SystemExit(f"cannot read {receiverConfigFile}")
    receiver = ReceiverGPSL1CA(receiverConfig, overwrite=True, gui=None)
    if "--fast" in sys.argv:
        receiver.run_fast()
    else:
        receiver.run()
    receiver.close()


if __name__ == "__main__":
    main()
