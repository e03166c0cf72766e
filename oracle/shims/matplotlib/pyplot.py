"""Empty stand-in for matplotlib.pyplot (imported unused by the reference)."""
