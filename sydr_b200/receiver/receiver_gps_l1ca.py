"""GPS L1 C/A receiver with the interface of sydr/receiver/receiver_gps_l1ca.py (ReceiverGPSL1CA):
reads [SATELLITES] include_prn and [CHANNELS] gps_l1ca from the receiver configuration, creates one
channel per PRN in the batched channel manager and requests tracking (receiver_gps_l1ca.py:44-89).
The channel class follows the channel configuration file: the Kaplan variant when its [TRACKING]
section carries the Kaplan keys (the reference selects it by editing an import,
receiver_gps_l1ca.py:17-18)."""
from __future__ import annotations

import configparser

from ..channel.channel_l1ca_borre import ChannelL1CA, ChannelStatusL1CA
from ..channel.channel_l1ca_kaplan import ChannelL1CA_Kaplan
from ..utils.enumerations import ChannelMessage
from .receiver import Receiver


class ReceiverGPSL1CA(Receiver):
    def __init__(self, configuration, overwrite=True, gui=None):
        channelConfig = configparser.ConfigParser()
        if not channelConfig.read(configuration['CHANNELS']['gps_l1ca']):
            raise FileNotFoundError(configuration['CHANNELS']['gps_l1ca'])
        kaplan = 'fll_bandwidth_pullin' in channelConfig['TRACKING']
        super().__init__(configuration, overwrite, gui, hostCopy=kaplan)
        self.prnList = list(map(int, self.configuration.get('SATELLITES', 'include_prn').split(',')))
        self.channelConfig = channelConfig
        self.channelClass = ChannelL1CA_Kaplan if kaplan else ChannelL1CA
        self.channelManager.addChannel(self.channelClass, channelConfig, len(self.prnList))
        for prn in self.prnList:
            channel = self.channelManager.requestTracking(prn)
            self.addChannelDatabase(channel)
            self.channelsStatus[channel.channelID] = ChannelStatusL1CA(channel.channelID, prn)

    def _processChannelResults(self, results: list):
        """receiver_gps_l1ca.py:93-134 without the satellite / ephemeris bookkeeping."""
        super()._processChannelResults(results)
        for packet in results:
            if packet is None:
                continue
            status = self.channelsStatus[packet['cid']]
            if packet['type'] == ChannelMessage.DECODING_UPDATE:
                status.subframeFlags[packet['subframe_id'] - 1] = True
            elif packet['type'] == ChannelMessage.CHANNEL_UPDATE:
                status.channelState = packet['state']
                status.trackFlags = packet['tracking_flags']
                status.tow = packet['tow']
                status.timeSinceTOW = packet['time_since_tow']
                status.unprocessedSamples = packet['unprocessed_samples']
                status.codeSinceTOW = packet['code_since_tow']
            elif packet['type'] not in (ChannelMessage.ACQUISITION_UPDATE, ChannelMessage.TRACKING_UPDATE):
                raise ValueError(f"Unknown channel message '{packet['type']}' received from channel {packet['cid']}.")

    def run_fast(self, chunk_seconds: float = 1.0):
        """Whole-file processing through the streaming path (Borre or Kaplan loop closure on the device): the PRNs of
        [SATELLITES] are searched in the first chunk, the ones found tracked to the end of the file (or
        ms_to_process), rows inserted column-wise.  Same database tables as run()."""
        from ..ingest import StreamingReceiver
        acq, trk = self.channelConfig['ACQUISITION'], dict(self.channelConfig['TRACKING'])
        rx = StreamingReceiver(self.rfSignal, self.prnList, len(self.prnList), chunk_seconds=chunk_seconds,
                               doppler_range=float(acq['doppler_range']), doppler_step=float(acq['doppler_steps']),
                               coh=int(acq['coherent_integration']), noncoh=int(acq['non_coherent_integration']),
                               threshold=float(acq['threshold']), channel_cfg=trk,
                               loop="kaplan" if self.channelClass is ChannelL1CA_Kaplan else "borre")
        try:
            ids = {int(ch.satelliteID): cid for cid, ch in self.channelManager.channels.items()}
            self.database.commit()                                  # the channel rows registered at construction
            out = rx.run_to_database(self.database, max_samples=self.msToProcess * self.rfSignal.samplesPerMs,
                                     channel_ids=ids)
        finally:
            rx.close()
        self.samplesCounter = self.msToProcess * self.rfSignal.samplesPerMs
        return out
