"""Stand-in for IPython (missing here); LearnerRecon imports display.clear_output only."""
