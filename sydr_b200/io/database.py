"""SQLite result sink with the interface and on-disk format of sydr/io/database.py
(DatabaseHandler), built for a receiver that produces results thousands of times faster than
real time.

Same tables and column-growing rule as the reference (database.py:118-213: `channel`,
`acquisition`, `tracking`, `decoding`, `position`, `measurement`, `gpsbrdc`; a key that is not
yet a column is added with the SQLite type of its first value, L78-93; lists and arrays are
pickled into BLOBs, L95-101; `ChannelMessage` values are dropped), same `addData` / `commit` /
`fetch*` / `close` methods, so a database written here is read by the reference's own
`fetchTracking` / `fetchTable` and vice versa.

What differs is the cost per row.  The reference builds one SQL string and calls `executemany`
once *per packet* (L71-107) - 720 000 statements per minute of signal for 12 channels.  Here
`commit()` inserts every run of packets that share a key set with one `executemany`, and
`addTrackingRecords()` takes the device's per-epoch record arrays column-wise (no per-epoch
dictionaries at all): one statement per channel and chunk.
"""
from __future__ import annotations

import logging
import os
import pickle
import sqlite3

import numpy as np

from ..utils.enumerations import ChannelMessage

# TRACKING_UPDATE packet keys in the order ChannelL1CA.runTracking fills them
# (sydr/channel/channel.py:198-200 `cid`, channel_l1ca_borre.py:432-449), then the three the
# receiver adds (sydr/receiver/receiver.py:357-360).
TRACKING_KEYS = ("cid", "i_early", "q_early", "i_prompt", "q_prompt", "i_late", "q_late", "dll", "pll", "fll",
                 "carrier_frequency", "code_frequency", "cn0", "pll_lock", "fll_lock", "lock_state",
                 "carrier_frequency_error", "code_frequency_error", "channel_id", "time", "time_sample")

# The Kaplan channel's TRACKING_UPDATE keys in its own order (channel_l1ca_kaplan.py:657-681).
KAPLAN_TRACKING_KEYS = ("cid", "i_early", "q_early", "i_prompt", "q_prompt", "i_late", "q_late", "carrier_frequency",
                        "code_frequency", "carrier_frequency_error", "code_frequency_error", "cn0", "pll_lock", "fll_lock",
                        "dll", "pll", "fll", "lock_state", "channel_id", "time", "time_sample")

_TABLES = {
    "channel": ("""CREATE TABLE IF NOT EXISTS channel (
                        id INTEGER PRIMARY KEY,
                        physical_id INTEGER,
                        system TEXT,
                        satellite_id INTEGER,
                        signal TEXT,
                        start_time FLOAT,
                        stop_time FLOAT,
                        start_sample INTEGER
                        );""",
                ["id", "physical_id", "system", "satellite_id", "signal", "start_time", "stop_time", "start_sample"]),
    "position": ("""CREATE TABLE IF NOT EXISTS position (
                        id INTEGER PRIMARY KEY,
                        time FLOAT,
                        time_sample INTEGER,
                        time_receiver TEXT,
                        x FLOAT,
                        y FLOAT,
                        z FLOAT,
                        clock FLOAT
                        );""",
                 ["id", "time", "time_sample", "time_receiver", "x", "y", "z", "clock"]),
    "measurement": ("""CREATE TABLE IF NOT EXISTS measurement (
                        id INTEGER PRIMARY KEY,
                        channel_id INTEGER,
                        time FLOAT,
                        time_sample FLOAT,
                        position_id INTEGER,
                        type FLOAT,
                        value FLOAT,
                        raw_value FLOAT,
                        residuals FLOAT,
                        FOREIGN KEY (channel_id) REFERENCES channels(id),
                        FOREIGN KEY (position_id) REFERENCES positions(id)
                        );""",
                    ["id", "channel_id", "time", "time_sample", "position_id", "type", "value", "raw_value", "residuals"]),
}
for _t in ("acquisition", "tracking", "decoding"):
    _TABLES[_t] = (f"""CREATE TABLE IF NOT EXISTS {_t} (
                        id INTEGER PRIMARY KEY,
                        channel_id INTEGER,
                        time FLOAT,
                        time_sample INTEGER,
                        FOREIGN KEY (channel_id) REFERENCES channels(id)
                        );""",
                   ["id", "channel_id", "time", "time_sample"])
_BRDC_COLUMNS = (("system_id", "TEXT"), ("satellite_id", "INTEGER"), ("datetime", "TEXT"), ("ura", "INTEGER"),
                 ("health", "INTEGER"), ("week", "INTEGER"), ("iode", "INTEGER"), ("iodc", "INTEGER"),
                 ("toe", "INTEGER"), ("toc", "INTEGER"), ("tgd", "FLOAT"), ("af0", "FLOAT"), ("af1", "FLOAT"),
                 ("af2", "FLOAT"), ("ecc", "FLOAT"), ("sqrtA", "FLOAT"), ("crs", "FLOAT"), ("deltan", "FLOAT"),
                 ("m0", "FLOAT"), ("cuc", "FLOAT"), ("cus", "FLOAT"), ("cic", "FLOAT"), ("omega0", "FLOAT"),
                 ("cis", "FLOAT"), ("i0", "FLOAT"), ("crc", "FLOAT"), ("omega", "FLOAT"), ("omegaDot", "FLOAT"),
                 ("iDot", "FLOAT"))


def _sql_type(val):
    """database.py:78-90 (bool is an int in Python, as there)."""
    if isinstance(val, int):
        return "INTEGER"
    if isinstance(val, float):
        return "FLOAT"
    if isinstance(val, str):
        return "TEXT"
    if isinstance(val, (list, np.ndarray)):
        return "BLOB"
    if isinstance(val, ChannelMessage):
        return None
    raise TypeError("Unknown type given in database.")


def cn0_column(first_epoch: int, n: int, sync_epoch: int) -> np.ndarray:
    """The `cn0` value of the Borre TRACKING_UPDATE packets of epochs [first_epoch, first_epoch + n):
    NaN, except 0.0 on every 20th epoch after bit synchronisation (channel_l1ca_borre.py:350,
    408-413: nbPrompt == LNAV_MS_PER_BIT with BIT_SYNC set)."""
    out = np.full(n, np.nan)
    if sync_epoch >= 0:
        k = np.arange(first_epoch, first_epoch + n)
        out[(k > sync_epoch) & ((k - sync_epoch) % 20 == 0)] = 0.0
    return out


class DatabaseHandler:
    def __init__(self, dbPath, overwrite=False):
        if overwrite and os.path.exists(dbPath):
            os.remove(dbPath)
        self.connection = sqlite3.connect(dbPath, detect_types=sqlite3.PARSE_DECLTYPES)
        self.cursor = self.connection.cursor()
        self.columns = {}
        self.dictBuffer = {}
        self._order = []               # (table, packet) in arrival order across tables
        self.sizeDictBuffer = 0
        self.maxSizeDictBuffer = 1000000
        self._initialise()
        logging.getLogger(__name__).info("Database initialized.")

    # ---- the reference's interface ---------------------------------------------------------
    def addData(self, table, data):
        """database.py:47-60: buffer one packet."""
        self.dictBuffer.setdefault(table, []).append(data)
        self.sizeDictBuffer += len(data)

    def commit(self):
        """database.py:64-113, with one INSERT statement per run of packets that share a key set."""
        logging.getLogger(__name__).info("Committing to database.")
        for table, inserts in self.dictBuffer.items():
            run_keys, run_rows = None, []
            for data in inserts:
                keys, values = [], []
                for key, val in data.items():
                    if key not in self.columns[table]:
                        mtype = _sql_type(val)
                        if mtype is None:
                            continue
                        if run_rows:                       # rows buffered so far predate the new column
                            self._insert(table, run_keys, run_rows)
                            run_keys, run_rows = None, []
                        self.addColumn(table, {key: mtype})
                        self.columns[table].append(key)
                    if isinstance(val, (list, np.ndarray)):
                        val = sqlite3.Binary(pickle.dumps(val, pickle.HIGHEST_PROTOCOL))
                    keys.append(key)
                    values.append(val)
                keys = tuple(keys)
                if keys != run_keys and run_rows:
                    self._insert(table, run_keys, run_rows)
                    run_rows = []
                run_keys = keys
                run_rows.append(values)
            if run_rows:
                self._insert(table, run_keys, run_rows)
        self.connection.commit()
        self.dictBuffer = {}
        self.sizeDictBuffer = 0

    def _insert(self, table, keys, rows):
        sqlstr = f"INSERT INTO {table} ({','.join(keys)}) VALUES ({','.join('?' * len(keys))});"
        self.cursor.executemany(sqlstr, rows)

    def addColumn(self, table, columnDict):
        """database.py:217-231."""
        for key, value in columnDict.items():
            self.cursor.execute(f"ALTER TABLE {table} ADD {key} {value}")

    def _initialise(self):
        """database.py:118-213."""
        for name in ("channel", "acquisition", "tracking", "decoding", "position", "measurement"):
            sql, cols = _TABLES[name]
            self.cursor.execute(sql)
            self.columns[name] = list(cols)
        cols = ",\n".join(f"{n} {t}" for n, t in _BRDC_COLUMNS)
        self.cursor.execute(f"CREATE TABLE IF NOT EXISTS gpsbrdc (id INTEGER PRIMARY KEY,\n{cols});")
        self.columns["gpsbrdc"] = ["id"] + [n for n, _ in _BRDC_COLUMNS]
        self.connection.commit()

    # ---- columnar fast path ----------------------------------------------------------------
    def addTrackingRecords(self, cid: int, records: np.ndarray, time, time_sample, cn0=None, channel_id=None,
                           fll: float = 0.0, kaplan: np.ndarray | None = None):
        """Insert the TRACKING_UPDATE rows of one channel straight from the device's per-epoch
        records (`sydr_trk_epoch` array): the same rows, in the same column order, that
        `addData("tracking", packet)` + `commit()` produce from ChannelL1CA.runTracking's packets
        (channel_l1ca_borre.py:432-449) after Receiver.addTrackingDatabase (receiver.py:357-362).
        `time` / `time_sample`: scalars or one value per epoch; `cn0`: see cn0_column()."""
        n = len(records)
        if n == 0:
            return
        if self.dictBuffer.get("tracking"):
            self.commit()                                       # keep arrival order
        if kaplan is not None:
            return self._addKaplanRecords(cid, records, kaplan, time, time_sample, channel_id)
        self._ensure_columns("tracking", TRACKING_KEYS, {"cid": "INTEGER", "lock_state": "INTEGER",
                                                         "channel_id": "INTEGER", "time_sample": "INTEGER"})
        corr = records["corr"]

        def col(x, conv=float):
            if np.ndim(x) == 0:
                return [conv(x)] * n
            return np.asarray(x).tolist() if conv is float else [int(v) for v in x]

        cols = [[int(cid)] * n] + [corr[:, k].tolist() for k in range(6)] + [
            records["dll"].tolist(), records["pll"].tolist(), [float(fll)] * n,
            records["carrier_freq"].tolist(), records["code_freq"].tolist(),
            (np.full(n, np.nan) if cn0 is None else np.asarray(cn0, dtype=np.float64)).tolist(),
            [0.0] * n, [0.0] * n, [0] * n,
            records["carrier_err"].tolist(), records["code_err"].tolist(),
            [int(cid if channel_id is None else channel_id)] * n, col(time), col(time_sample, int)]
        self._insert("tracking", TRACKING_KEYS, zip(*cols))

    def _addKaplanRecords(self, cid, records, kaplan, time, time_sample, channel_id):
        """Rows of the Kaplan channel's packets (channel_l1ca_kaplan.py:657-681) from the device records and
        their Kaplan extras: `dll` / `pll` / `fll` are the discriminators, `carrier_frequency_error` /
        `code_frequency_error` the loop-filter outputs, `lock_state` the LoopLockState value."""
        n = len(records)
        assert len(kaplan) == n
        self._ensure_columns("tracking", KAPLAN_TRACKING_KEYS, {"cid": "INTEGER", "lock_state": "INTEGER",
                                                                "channel_id": "INTEGER", "time_sample": "INTEGER"})
        corr = records["corr"]

        def col(x, conv=float):
            if np.ndim(x) == 0:
                return [conv(x)] * n
            return np.asarray(x).tolist() if conv is float else [int(v) for v in x]

        cols = [[int(cid)] * n] + [corr[:, k].tolist() for k in range(6)] + [
            records["carrier_freq"].tolist(), records["code_freq"].tolist(), records["pll"].tolist(),
            records["dll"].tolist(), kaplan["cn0"].tolist(), kaplan["pll_lock"].tolist(), kaplan["fll_lock"].tolist(),
            records["code_err"].tolist(), records["carrier_err"].tolist(), kaplan["fll"].tolist(),
            [int(v) for v in kaplan["lock_state"]],
            [int(cid if channel_id is None else channel_id)] * n, col(time), col(time_sample, int)]
        self._insert("tracking", KAPLAN_TRACKING_KEYS, zip(*cols))

    def _ensure_columns(self, table, keys, int_keys):
        for key in keys:
            if key not in self.columns[table]:
                self.addColumn(table, {key: int_keys.get(key, "FLOAT")})
                self.columns[table].append(key)

    # ---- queries (database.py:381-470) -------------------------------------------------------
    def _select(self, table, channelID=None, extra=""):
        where = "" if channelID is None else f" WHERE channel_id={channelID}{extra}"
        return self._unpackData(self.cursor.execute(f"SELECT * FROM {table}{where};").fetchall())

    def fetchTracking(self, channelID=None):
        return self._select("tracking", channelID)

    def fetchAcquisition(self, channelID=None):
        return self._select("acquisition", channelID)

    def fetchMeasurements(self, channelID=None, mtype=None):
        return self._select("measurement", channelID, f" AND type='{mtype}'")

    def fetchTable(self, tableName):
        rows = self.cursor.execute(f"SELECT * FROM {tableName};").fetchall()
        names = [d[0] for d in self.cursor.description]
        return [dict(zip(names, r)) for r in rows]

    def sqlRequest(self, request):
        return self._unpackData(self.cursor.execute(request).fetchall())

    def _unpackData(self, fetchedData):
        names = [d[0] for d in self.cursor.description]
        return [{n: (pickle.loads(v) if isinstance(v, bytes) else v) for n, v in zip(names, row)}
                for row in fetchedData]

    def close(self):
        self.commit()
        self.connection.close()
