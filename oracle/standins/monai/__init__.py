"""ORACLE / TEST INFRASTRUCTURE ONLY. Stand-in package so /root/reference/model/loss.py:4 imports."""
