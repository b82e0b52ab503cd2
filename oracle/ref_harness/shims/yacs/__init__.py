"""Import shim for the golden generators: the reference imports ``yacs.config.CfgNode``; yacs is not installed here."""
