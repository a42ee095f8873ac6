"""Post-processing tools (mirror of ``mlmc/tool`` for the estimation path)."""
