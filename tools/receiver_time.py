"""main.py's two modes on a BASELINE configs[0]-like recording (4 MS/s int8, 8 PRNs, 3 s): the reference's
per-millisecond loop over the batched ChannelManager vs the streaming path; both into SQLite."""
import configparser, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sydr_b200 import synth
from sydr_b200.receiver.receiver_gps_l1ca import ReceiverGPSL1CA
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fs, nbits, sec = 4e6, 8, 3.0
tmp = tempfile.mkdtemp(prefix="sydr_rx_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
sc = synth.make_scenario(fs, nbits, sec + 0.13, synth.PRNS_8, 1001, 100.0)
path = os.path.join(tmp, "rec.bin")
synth.write_file(path, synth.generate_iq(sc))
for ini in ("channel_GPS_L1CA_borre.ini", "channel_GPS_L1CA_kaplan.ini"):
    for mode in ("run", "run_fast"):
        cfg = configparser.ConfigParser()
        cfg.read(os.path.join(ROOT, "config", "receiver.ini"))
        cfg["DEFAULT"].update({"name": f"{ini[:-4]}_{mode}", "ms_to_process": str(int(sec * 1000)), "outfolder": tmp})
        cfg["RFSIGNAL"].update({"filepath": path, "sampling_frequency": str(fs), "data_size": str(nbits)})
        cfg["SATELLITES"]["include_prn"] = ",".join(str(p) for p in synth.PRNS_8)
        cfg["CHANNELS"]["gps_l1ca"] = os.path.join(ROOT, "config", "channels", ini)
        rx = ReceiverGPSL1CA(cfg, overwrite=True)
        t0 = time.perf_counter()
        getattr(rx, mode)()
        rx.database.commit()
        dt = time.perf_counter() - t0
        rows = len(rx.database.fetchTable("tracking"))
        rx.close()
        print(f"{ini:32s} {mode:9s}: {dt:7.2f} s for {sec:g} s of signal = RTF {sec / dt:7.2f}, {rows} tracking rows")
