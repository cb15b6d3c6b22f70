"""empty stand-in (plotting is out of scope for goldens)"""
