"""Receiver with the interface of sydr/receiver/receiver.py (Receiver): what `main.py` drives.

Same configuration object (the `configparser` content of config/receiver.ini: [DEFAULT] name,
ms_to_process, outfolder; [RFSIGNAL]; [MEASUREMENTS]), same attributes (`name`, `msToProcess`,
`outfolder`, `rfSignal`, `database`, `channelManager`, `samplesCounter`, `channelsStatus`,
`receiverState`) and methods (`run`, `close`, `_processChannelResults`,
`_updateDatabaseFromChannels`, `add{Acquisition,Tracking,Decoding,Channel}Database`).

`run()` is the reference's loop (receiver.py:120-143) - one millisecond from the file, one
`ChannelManager.run()` tick, packets into the database - except that the channel manager behind it
is the batched GPU dispatcher instead of one process per channel.  `run_fast()` is the whole-file
path for when nobody needs the per-tick packet schedule: `StreamingReceiver` (file -> pinned
buffers -> GPU -> column-wise database inserts), the same tables in the same SQLite format.

Out of scope here: measurements, position (`computeGNSSMeasurements`,
`LeastSquareEstimation`), the enlighten GUI and the HTML report; a GUI object is accepted and
ignored unless it has the methods the reference calls.
"""
from __future__ import annotations

import logging
import os
import time

from ..channel.channelManager import ChannelManager
from ..io.database import DatabaseHandler
from ..signal.rfsignal import RFSignal
from ..utils.enumerations import ChannelMessage


class ReceiverState:
    OFF, IDLE, INIT, NAVIGATION = "OFF", "IDLE", "INIT", "NAVIGATION"      # sydr/utils/enumerations.py:77-81


class Receiver:
    def __init__(self, configuration, overwrite=True, gui=None, hostCopy=False):
        """receiver.py:58-97.  hostCopy: keep a host copy of the 100 ms ring (channel classes that are ticked
        stand-alone read it)."""
        self.configuration = configuration
        self.name = str(configuration['DEFAULT']['name'])
        self.msToProcess = int(configuration['DEFAULT']['ms_to_process'])
        self.outfolder = str(configuration['DEFAULT']['outfolder'])
        os.makedirs(self.outfolder, exist_ok=True)
        self.rfSignal = RFSignal(configuration['RFSIGNAL'])
        self.database = DatabaseHandler(f"{self.outfolder}/{self.name}.db", overwrite)
        self.measurementFrequency = (float(configuration['MEASUREMENTS']['frequency'])
                                     if 'MEASUREMENTS' in configuration else 1.0)
        self.receiverState = ReceiverState.IDLE
        self.channelManager = ChannelManager(self.rfSignal, keepCorrelationMaps=False, hostCopy=hostCopy)
        self.samplesCounter = 0
        self.channelsStatus = {}
        self.satelliteDict = {}
        self.gui = gui

    # ---- the reference's loop -----------------------------------------------------------------
    def run(self):
        """receiver.py:101-143: one millisecond per iteration."""
        logging.getLogger(__name__).info(f"Processing in receiver {self.name} started.")
        self.receiverState = ReceiverState.INIT
        msPerLoop = 1
        for _ in range(self.msToProcess):
            data = self.rfSignal.getMilliseconds(nbMilliseconds=msPerLoop)
            if len(data) < msPerLoop * self.rfSignal.samplesPerMs:
                break                                                       # end of file
            self.channelManager.addNewRFData(data)
            self.samplesCounter += msPerLoop * self.rfSignal.samplesPerMs
            results = self.channelManager.run()
            self._processChannelResults(results)
            self.computeGNSSMeasurements()
            self._updateGUI()
        self.database.commit()

    def computeGNSSMeasurements(self):
        """Measurements and position are outside the hot-path scope."""
        return

    def _updateGUI(self):
        if self.gui is not None and hasattr(self.gui, "updateReceiverGUI"):
            self.gui.updateReceiverGUI(self)

    def _processChannelResults(self, results: list):
        """receiver.py:147-164."""
        self._updateDatabaseFromChannels(results)

    def _updateDatabaseFromChannels(self, results: list):
        """receiver.py:276-301."""
        for packet in results:
            if packet is None:
                continue
            if packet['type'] == ChannelMessage.ACQUISITION_UPDATE:
                self.addAcquisitionDatabase(packet)
            elif packet['type'] == ChannelMessage.TRACKING_UPDATE:
                self.addTrackingDatabase(packet)
            elif packet['type'] == ChannelMessage.DECODING_UPDATE:
                self.addDecodingDatabase(packet)

    def _stamp(self, result: dict):
        channel = self.channelManager.getChannel(result['cid'])
        for key in [k for k, v in result.items() if v is None]:        # e.g. the correlation map when it is not kept
            del result[key]
        result["channel_id"] = channel.channelID
        result["time"] = time.time()
        result["time_sample"] = self.samplesCounter
        return result

    def addAcquisitionDatabase(self, result: dict):
        """receiver.py:305-333."""
        self.database.addData("acquisition", self._stamp(result))

    def addTrackingDatabase(self, result: dict):
        """receiver.py:337-366 (lock_state is an enumeration in the Kaplan packets: stored by value)."""
        if "lock_state" in result and not isinstance(result["lock_state"], (int, float)):
            result["lock_state"] = int(result["lock_state"])
        self.database.addData("tracking", self._stamp(result))

    def addDecodingDatabase(self, result: dict):
        """receiver.py:370-400."""
        self.database.addData("decoding", self._stamp(result))

    def addChannelDatabase(self, channel):
        """receiver.py:404-430 (enumerations stored by name: SQLite has no column type for them)."""
        self.database.addData("channel", {
            "id": channel.channelID, "system": str(channel.systemID), "satellite_id": int(channel.satelliteID),
            "signal": str(channel.signalID), "start_time": time.time(), "start_sample": self.samplesCounter})

    def close(self):
        """receiver.py:484-499."""
        self.channelManager.close()
        self.database.close()
        try:
            self.rfSignal.closeFile()
        except Warning:
            pass
