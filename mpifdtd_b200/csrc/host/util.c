/* util.c -- small host utilities the reference's driver and viewer link against
 * (function.c:1-81 and myComplex.c:1-51 of rennone/mpiFDTD): zero-filled
 * allocators, squared magnitude, bilinear samplers over host mirrors, and the
 * output-directory helpers.  Error convention as upstream: printf + exit(2). */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include "mpifdtd_plugin.h"

double *newDouble(int size)        { return (double *)calloc((size_t)(size > 0 ? size : 1), sizeof(double)); }
dcomplex *newDComplex(int size)    { return (dcomplex *)calloc((size_t)(size > 0 ? size : 1), sizeof(dcomplex)); }
void freeDouble(double *array)     { free(array); }
void freeDComplex(dcomplex *array) { free(array); }

/* myComplex.c:34-38 -- note: SQUARED magnitude */
double cnorm(dcomplex c)
{
  double re = creal(c), im = cimag(c);
  return re * re + im * im;
}

/* 4-point bilinear sample of a [width][height] array stored k = i*height + j
 * (myComplex.c:40-51, function.c:12-24); weights multiplied in the same order. */
double complex cbilinear(dcomplex *p, double x, double y, int width, int height)
{
  (void)width;
  int i = floor(x), j = floor(y);
  double fx = x - i, fy = y - j;
  const dcomplex *q = p + (i * height + j);
  return q[0] * (1.0 - fx) * (1.0 - fy) + q[height] * fx * (1.0 - fy)
       + q[1] * (1.0 - fx) * fy         + q[height + 1] * fx * fy;
}

double dbilinear(double *p, double x, double y, int width, int height)
{
  (void)width;
  int i = floor(x), j = floor(y);
  double fx = x - i, fy = y - j;
  const double *q = p + (i * height + j);
  return q[0] * (1.0 - fx) * (1.0 - fy) + q[height] * fx * (1.0 - fy)
       + q[1] * (1.0 - fx) * fy         + q[height + 1] * fx * fy;
}

FILE *FileOpen(const char *file_name, const char *mode)
{
  FILE *fp = fopen(file_name, mode);
  if (fp == NULL) { printf("cannot open file %s \n", file_name); exit(2); }
  return fp;
}
FILE *openFile(const char *file_name) { return FileOpen(file_name, "w"); }

/* function.c:51-80: mkdir 0775-ish, chdir or die */
bool makeDirectory(const char *name)
{
  return mkdir(name, S_IRWXU | S_IRWXG | S_IROTH | S_IXOTH) == 0;
}
void moveDirectory(const char *name)
{
  if (chdir(name) != 0) { printf("cannot move to %s\n", name); exit(2); }
  printf("move to %s\n", name);
}
void makeAndMoveDirectory(const char *name)
{
  makeDirectory(name);
  moveDirectory(name);
}
