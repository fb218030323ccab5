#!/bin/bash
python tools/pad_probe.py
