"""Empty ``matplotlib.pyplot`` stand-in (run the reference with save_visualizations=False)."""
