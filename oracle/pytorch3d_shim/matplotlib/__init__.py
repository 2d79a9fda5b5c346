"""Empty ``matplotlib`` stand-in so the reference's ``mhmocap/predict.py:6`` imports (TEST INFRASTRUCTURE ONLY)."""
