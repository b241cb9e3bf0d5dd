"""Stand-in for ``fvcore`` (absent here): only ``fvcore.nn.FlopCountAnalysis`` as the reference uses it."""
