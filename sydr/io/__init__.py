"""`sydr.io`: `sydr.io.database` is sydr_b200.io.database (alias, sydr/__init__.py); the HTML report is a stand-in."""
