"""Test infrastructure only: CPU restatement of the reference path (see oracle/uav_oracle.py). Never imported by the product."""
