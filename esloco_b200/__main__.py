from .esloco import main

main()
