"""The `sydr` import surface (sydr/__init__.py): the reference's module paths resolve to the sydr_b200 modules, and the
statements of the reference's main.py (main.py:4-41) execute against this repository."""
import configparser
import importlib
import os

import numpy as np
import pytest

import helpers as H  # noqa: F401

PATHS = ["sydr.dsp.acquisition", "sydr.dsp.tracking", "sydr.dsp.decoding", "sydr.dsp.lockindicator", "sydr.signal.gnsssignal",
         "sydr.signal.rfsignal", "sydr.channel.channel", "sydr.channel.channelManager", "sydr.channel.channel_l1ca_borre",
         "sydr.channel.channel_l1ca_kaplan", "sydr.receiver.receiver", "sydr.receiver.receiver_gps_l1ca", "sydr.io.database",
         "sydr.utils.circularbuffer", "sydr.utils.constants", "sydr.utils.enumerations",
         "sydr.old.acquisition.acquisition_pcps_c", "sydr.old.tracking.tracking_epl_c"]


@pytest.mark.parametrize("path", PATHS)
def test_reference_module_paths_resolve_to_the_same_modules(path):
    a = importlib.import_module(path)
    b = importlib.import_module("sydr_b200." + path[len("sydr."):])
    assert a is b


def test_names_the_reference_callers_import():
    from sydr.channel.channel_l1ca_borre import ChannelL1CA, ChannelStatusL1CA  # noqa: F401
    from sydr.dsp.acquisition import PCPS, TwoCorrelationPeakComparison  # noqa: F401
    from sydr.dsp.tracking import EPL, DLL_NNEML, PLL_costa, BorreLoopFilter, LoopFiltersCoefficients  # noqa: F401
    from sydr.io.database import DatabaseHandler  # noqa: F401
    from sydr.receiver.receiver_gps_l1ca import ReceiverGPSL1CA  # noqa: F401
    from sydr.signal.gnsssignal import GenerateGPSGoldCode, UpsampleCode  # noqa: F401
    from sydr.signal.rfsignal import RFSignal  # noqa: F401
    from sydr.utils.circularbuffer import CircularBuffer  # noqa: F401
    from sydr.utils.enumerations import ChannelMessage, TrackingFlags  # noqa: F401
    with pytest.raises(ImportError):
        importlib.import_module("sydr.navigation.lse")           # out of scope stays absent


@pytest.mark.gpu
def test_reference_main_flow(tmp_path, monkeypatch):
    """The statements of the reference's main.py, one for one, with its own imports, on a synthetic 4 MS/s int8 file:
    configuration -> GUI -> logger -> ReceiverGPSL1CA(overwrite=True, gui=gui) -> run -> close -> Visualisation.run."""
    import sqlite3
    from sydr_b200 import synth
    sc = synth.make_scenario(4e6, 8, 0.45, (3, 7), 77, 250.0)     # 300 ms to process + the reader's 120 ms chunk margin
    iq_path = tmp_path / "iq.bin"
    synth.write_file(str(iq_path), synth.generate_iq(sc))
    chan_ini = tmp_path / "channel.ini"
    src = configparser.ConfigParser()
    assert src.read(os.path.join(H.ROOT, "config", "receiver.ini"))
    base_chan = configparser.ConfigParser()
    assert base_chan.read(os.path.join(H.ROOT, src["CHANNELS"]["gps_l1ca"]))
    with open(chan_ini, "w") as f:
        base_chan.write(f)
    src["RFSIGNAL"]["filepath"] = str(iq_path)
    src["RFSIGNAL"]["sampling_frequency"] = "4e6"
    src["RFSIGNAL"]["data_size"] = "8"
    src["DEFAULT"]["name"] = "MAINFLOW"
    src["DEFAULT"]["ms_to_process"] = "300"
    src["DEFAULT"]["outfolder"] = str(tmp_path)
    src["SATELLITES"]["include_prn"] = "3,7"
    src["CHANNELS"]["gps_l1ca"] = str(chan_ini)
    os.makedirs(tmp_path / "config")
    with open(tmp_path / "config" / "receiver.ini", "w") as f:
        src.write(f)
    monkeypatch.chdir(tmp_path)

    # ---- main.py:4-8
    from sydr.enlightengui import EnlightenGUI
    from sydr.receiver.receiver_gps_l1ca import ReceiverGPSL1CA
    from sydr.io.visualisation import Visualisation
    import sydr.logger as logger
    # ---- main.py:15-41
    receiverConfigFile = './config/receiver.ini'
    receiverConfig = configparser.ConfigParser()
    receiverConfig.read(receiverConfigFile)
    gui = EnlightenGUI()
    gui.updateMainStatus(stage='Initialize', status='RUNNING')
    logger.configureLogger(name=__name__, filepath='./config/logging.ini')
    receiver = ReceiverGPSL1CA(receiverConfig, overwrite=True, gui=gui)
    receiver.run()
    receiver.close()
    gui.updateMainStatus(stage='Create report', status='RUNNING')
    visual = Visualisation(receiverConfig)
    visual.run()
    gui.updateMainStatus(stage='PROCESSING COMPLETED', status='DONE')

    # the run left the reference's database behind: both satellites acquired and tracked
    assert os.path.exists(tmp_path / "MAINFLOW.db"), os.listdir(tmp_path)
    con = sqlite3.connect(str(tmp_path / "MAINFLOW.db"))
    n_trk = con.execute("select count(*) from tracking").fetchone()[0]
    n_acq = con.execute("select count(*) from acquisition").fetchone()[0]
    con.close()
    assert n_acq == 2 and n_trk >= 2 * 280
    assert (gui.stage, gui.status) == ("PROCESSING COMPLETED", "DONE")
