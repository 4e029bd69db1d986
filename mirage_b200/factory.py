"""Decorator registry with the reference's interface (mutils/factory.py:5-36).

``add, table = get_factory_adder()``; ``@add`` registers under the object's ``__name__``,
``@add('key')`` under an explicit key; ``table[key]`` returns the registered callable.
"""
from __future__ import annotations


def get_factory_adder():
    table: dict = {}

    def add(arg):
        if callable(arg) and not isinstance(arg, str):
            table[arg.__name__] = arg
            return arg

        def deco(obj):
            table[arg] = obj
            return obj
        return deco

    return add, table
